"""SURVEY 8f N3 -- the on-disk store reader, date-indexed datasets, transform fitting and the batch loader (host side) against
fixtures produced by the REAL reference DataHandler on the same synthetic store (tests/golden/store.npz <-
``python -m oracle.make_golden store``; the store itself is regenerated from oracle/store.py).  No GPU: the loader runs with
``device=None`` and yields host tensors; the device-resident path is tests/test_store_gpu.py."""
import json
import os

import numpy as np
import pytest
import torch

import wsr
from conftest import load_golden
from oracle import store

TOL = 2e-5


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    return store.write_store(str(tmp_path_factory.mktemp("store")))


@pytest.fixture(scope="module")
def handler(root):
    builder, transforms = wsr.sub("data.dataset_builder"), wsr.sub("data.transforms")
    s = store.SPEC
    dh = builder.DataHandler(root, list(store.VARIABLES), root, s["months_subset"], s["groups"], transforms.GlobalStandardScaling,
                             s["train"][0], s["train"][1], s["val"][0], s["val"][1], s["val_batch_size"], s["train_batch_size"], False, 2,
                             device=None)
    dh.process_data()
    return dh


def _close(a, b, tol=TOL):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm()) < tol


def test_reader_index_and_items(root):
    R = wsr.sub("data.npy_reader").WNPYReader
    r = R(os.path.join(root, "hr", "t2m"))
    assert len(r) == store.HOURS and r.is_time_variate() and r.get_channel_count() == 1 and r.name == "t2m"
    stamps = r.get_valid_time_stamps()
    assert stamps[0] == np.datetime64("2000-01-01T00") and stamps[-1] == np.datetime64("2000-03-05T23")
    x = r[np.datetime64("2000-02-01T05")]
    assert x.shape == (1, 1, 32, 64) and torch.equal(x[0, 0], torch.from_numpy(store.field(0, 31 * 24 + 5)))
    assert torch.equal(r[7], r[stamps[7]])
    # bulk path: same bytes as the per-sample path, also into a channel slice of a wider staging buffer
    buf = np.zeros((5, 3, 32, 64), dtype=np.float32)
    items = [stamps[3], stamps[100], stamps[4], stamps[1559], stamps[0]]
    r.read_into(buf[:, 1:2], items)
    for i, t in enumerate(items):
        assert np.array_equal(buf[i, 1], r[t][0, 0].numpy())
    assert not buf[:, 0].any() and not buf[:, 2].any()
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(4) as pool:
        again = np.empty((5, 1, 32, 64), dtype=np.float32)
        r.read_into(again, [3, 100, 4, 1559, 0], pool)
    assert np.array_equal(again[:, 0], buf[:, 1])


def test_reader_rejects_bad_layouts(root, tmp_path):
    R = wsr.sub("data.npy_reader").WNPYReader
    with pytest.raises(AssertionError):
        R(str(tmp_path / "missing"))
    bad = tmp_path / "bad"
    (bad / "meta").mkdir(parents=True)
    with pytest.raises(AssertionError):          # no samples directory
        R(str(bad))
    # a gap in the hourly series
    gap = tmp_path / "gap"
    (gap / "meta").mkdir(parents=True)
    (gap / "samples" / "2000").mkdir(parents=True)
    with open(os.path.join(root, "lr", "t2m", "meta", "metadata.json")) as fh:
        meta = json.load(fh)
    with open(gap / "meta" / "metadata.json", "w") as fh:
        json.dump(meta, fh)
    for name in ("2000-01-01-00", "2000-01-01-01", "2000-01-01-03"):
        np.save(gap / "samples" / "2000" / (name + ".npy"), np.zeros((8, 16), np.float32))
    np.save(gap / "samples" / "2000" / "notes.npy.bak.npy", np.zeros(1))        # ignored: not <date>.npy
    with pytest.raises(AssertionError, match="missing data"):
        R(str(gap))


def test_time_variate_ranges(root):
    R = wsr.sub("data.npy_reader").WNPYReader
    D = wsr.sub("data.datasets")
    r = R(os.path.join(root, "lr", "z500"))
    d = D.TimeVariateData(r, name="x", lead_time=0, min_date="2000-01-01-00", max_date="2000-02-01-00")
    assert len(d) == 744 and d.name == "x"
    d.add_data_by_date("2000-03-01-00", "2000-03-04-00")
    assert len(d) == 816 and d.max_date == np.datetime64("2000-03-04T00") and len(d.date_ranges) == 2
    with pytest.raises(AssertionError, match="overlaps"):
        d.add_data_by_date("2000-01-31-00", "2000-02-02-00")
    with pytest.raises(AssertionError, match="beyond the range"):
        d.add_data_by_date("2000-03-05-00", "2000-04-05-00")
    with pytest.raises(AssertionError, match="beyond the range"):
        D.TimeVariateData(r, min_date="1999-12-31-00")
    with pytest.raises(Exception, match="invalid date string"):
        D.TimeVariateData(r, min_date="2000/01/01")
    with pytest.raises(AssertionError, match="must be earlier"):
        D.TimeVariateData(r, min_date="2000-02-01-00", max_date="2000-01-01-00")
    t, name, month = d[744]
    assert month == 0 and name == "x" and torch.equal(t, r[np.datetime64("2000-03-01T00")])        # no transform fitted -> key 0
    assert d.months_of([0, 744]) == [0, 0]
    assert list(d.stamps_of([0, 743, 744])) == [np.datetime64("2000-01-01T00"), np.datetime64("2000-01-31T23"), np.datetime64("2000-03-01T00")]
    # lead time and delays
    dl = D.TimeVariateData(r, lead_time=2, delays=[-1], min_date="2000-01-02-00", max_date="2000-01-03-00")
    items = dl[0]
    assert len(items) == 2 and dl.get_channel_count() == 2
    # delays become [0, -1] (0 is always prepended): the item at 00h + 2h lead, then the one an hour before it
    assert torch.equal(items[0][0], r[np.datetime64("2000-01-02T02")]) and torch.equal(items[1][0], r[np.datetime64("2000-01-02T01")])
    assert d.summarize()["number_of_intervals"] == 2 and d.summarize()["date_range"] == ["2000-01-01-00", "2000-03-04-00"]


def test_handler_matches_reference(handler):
    g = load_golden("store")
    train_set, val_set = handler.get_datasets()
    assert len(train_set) == int(g["train_len"]) and len(val_set) == int(g["val_len"])
    assert len(handler.train_loader) == int(g["train_batches"]) and len(handler.val_loader) == int(g["val_batches"])
    first = list(train_set.data_groups["lr"].values())[0]
    stamps = first.stamps_of(g["probe"]).astype("datetime64[h]").astype(np.int64)
    assert np.array_equal(stamps, g["probe_stamps"])
    td = handler.get_data_transformer().transformation_dict
    for v in store.VARIABLES:
        for kind in ("lr", "hr"):
            assert sorted(td[v][kind]) == [1, 3]
            for month, t in td[v][kind].items():
                assert float(t._mean) == pytest.approx(float(g["mean.%s.%s.%d" % (v, kind, month)]), rel=1e-6)
                assert float(t._std()) == pytest.approx(float(g["std.%s.%s.%d" % (v, kind, month)]), rel=1e-5)
    assert int(train_set.get_channel_count("lr")) == int(g["channels.lr"])
    assert np.allclose(handler.get_metadata().hr_lat, g["meta.hr_lat"])
    assert train_set.get_data_names() == {"lr": ("lr_t2m", "lr_z500"), "hr": ("hr_t2m", "hr_z500")}
    # one training item through the reference's per-sample interface
    item = train_set[800]
    assert item[0][0][2] == int(g["item800.month"]) and item[0][1][1] == "lr_z500"
    assert _close(torch.cat([v[0] for v in item[0]], 1), g["item800.lr"]) and _close(torch.cat([v[0] for v in item[1]], 1), g["item800.hr"])
    # first validation batch of the staged loader (raw reads -> batch statistics -> bicubic)
    batch, months = next(iter(handler.val_loader))
    assert months == [int(m) for m in g["val0.months"]]
    for k in ("HR", "LR", "SR"):
        assert batch[k].shape == g["val0." + k].shape and _close(batch[k], g["val0." + k]), k
    # by date, and back to physical units
    by_date, bm = handler.get_data_by_date("2000-03-05-07")
    assert bm == [int(m) for m in g["date.months"]]
    for k in ("HR", "LR", "SR"):
        assert _close(by_date[k], g["date." + k]), k
    inv = handler.get_data_transformer().inverse_transform(by_date, bm)
    for k in ("HR", "LR", "SR"):
        assert _close(inv[k], g["date_inv." + k], 1e-6), k
    with pytest.raises(AssertionError, match="beyond the range"):
        handler.get_data_by_date("2000-01-05-07")


def test_loader_epoch_order_and_shuffle(handler):
    L = wsr.sub("data.dataset_builder").DeviceBatchLoader
    train_set, val_set = handler.get_datasets()
    seq = L(val_set, 5, shuffle=False, num_workers=2, device=None)
    batches = list(seq)
    assert len(batches) == len(seq) == 48 // 5                       # drop_last
    per_sample = wsr.sub("data.dataset_builder").form_batch([val_set[i] for i in range(5, 10)])
    for k in ("HR", "LR", "SR"):
        assert _close(batches[1][0][k], per_sample[0][k], 1e-6), k
    sh = L(train_set, 16, shuffle=True, num_workers=2, device=None, seed=3)
    a = [b[0]["LR"][:, 0, 0, 0] for b in sh]
    b = [b[0]["LR"][:, 0, 0, 0] for b in sh]
    assert len(a) == 816 // 16 and not torch.equal(torch.cat(a), torch.cat(b))          # a fresh permutation per epoch
    months = [m for bt in L(train_set, 16, shuffle=True, device=None, seed=4) for m in bt[1]]
    assert set(months) == {1, 3}


def test_fit_paths_agree(root):
    """Bulk (read_into + float64) and per-sample (reference update rule in fp32) fitting give the same statistics; the Local
    variant keeps per-grid-point maps."""
    R = wsr.sub("data.npy_reader").WNPYReader
    D, T = wsr.sub("data.datasets"), wsr.sub("data.transforms")
    r = R(os.path.join(root, "lr", "t2m"))
    d = D.TimeVariateData(r, lead_time=0, min_date="2000-01-01-00", max_date="2000-01-11-00")
    bulk = T.GlobalStandardScaling().fit(d, batch_size=64)
    ref = T.GlobalStandardScaling()
    for chunk in d.get_batch(range(len(d)), chunk_size=100):
        ref._update_parameters(chunk)
    assert bulk._count == ref._count == 240 * 8 * 16
    assert float(bulk._mean) == pytest.approx(float(ref._mean), rel=1e-6) and float(bulk._std()) == pytest.approx(float(ref._std()), rel=1e-5)
    loc = T.LocalStandardScaling().fit(d)
    assert loc._mean.shape == (1, 1, 8, 16)
    x = d[5][0]
    assert _close(loc.revert(loc.transform(x)), x, 1e-6)
    with pytest.raises(Exception, match="only be called once"):
        bulk.fit(d)
    assert T.get_transformation_by_name("IdentityTransform")().transform(x) is x
    with pytest.raises(Exception, match="Unknown transformation"):
        T.get_transformation_by_name("MinMax")


def test_month_windows_and_helpers():
    U = wsr.sub("data.utils")
    w = [(U.date_to_str(a), U.date_to_str(b)) for a, b in U.month_windows("2000-01-10-00", "2000-03-05-00")]
    # the reference's walk: the first window is one calendar month long from min_date, later ones are cut at month starts
    assert w == [("2000-01-10-00", "2000-02-10-00"), ("2000-02-10-00", "2000-03-01-00"), ("2000-03-01-00", "2000-03-05-00")]
    assert U.find_group_idx(3, [[1, 2], [3]]) == 2 and U.find_group_idx(5, [[1, 2], [3]]) is None
    assert U.is_full_year(None) and U.is_full_year(list(range(1, 13))) and not U.is_full_year([1])
    U.validate_group_months_subset([1, 3], [[1], [3]])
    with pytest.raises(AssertionError):
        U.validate_group_months_subset([1, 3], [[1], [2]])
    with pytest.raises(AssertionError):
        U.validate_month_subset([0, 13])
    assert U.get_month_idx("2000-07-01-00") == 7


@pytest.mark.parametrize("name", ["LocalStandardScaling", "IdentityTransform"])
def test_loader_per_sample_fallback_transforms(root, name):
    """Transforms the staged path does not cover (per-grid-point statistics, identity): the loader falls back to the reference's
    per-sample interface and still yields the collate contract."""
    builder, transforms = wsr.sub("data.dataset_builder"), wsr.sub("data.transforms")
    dh = builder.DataHandler(root, ["t2m"], root, [1], [[1]], transforms.get_transformation_by_name(name), "2000-01-01-00", "2000-01-06-00",
                             "2000-01-06-00", "2000-01-07-00", 4, 4, False, 1, device=None)
    dh.process_data()
    loader = dh.val_loader
    assert loader._bulk == (name == "IdentityTransform")
    batch, months = next(iter(loader))
    assert batch["LR"].shape == (4, 1, 8, 16) and batch["HR"].shape == (4, 1, 32, 64) and batch["SR"].shape == (4, 1, 32, 64)
    assert months == [1, 1, 1, 1]
    raw = dh.val_dataset.data_groups["hr"]["hr_t2m"].wnpy_reader[np.datetime64("2000-01-06T00")]
    if name == "IdentityTransform":
        assert torch.equal(batch["HR"][:1], raw)
    else:
        t = dh.get_data_transformer().get_transform("t2m", "hr")[1]
        assert t._mean.shape == (1, 1, 32, 64) and _close(batch["HR"][:1], (raw - t._mean) / t._std(), 1e-6)
        inv = dh.get_data_transformer().inverse_transform({"HR": batch["HR"]}, months)
        assert _close(inv["HR"][:1], raw, 1e-6)


def test_loader_rank_shards_partition_the_global_batches(handler):
    """One process per GPU: ranks walk the same (shuffled) batch sequence and each reads only its contiguous slice."""
    L = wsr.sub("data.dataset_builder").DeviceBatchLoader
    train_set, _ = handler.get_datasets()
    full = list(L(train_set, 8, shuffle=True, num_workers=1, device=None, seed=5))
    parts = [list(L(train_set, 8, shuffle=True, num_workers=1, device=None, seed=5, shard=(r, 4))) for r in range(4)]
    assert all(len(p) == len(full) == 816 // 8 for p in parts)
    for i in (0, 7, len(full) - 1):
        for k in ("HR", "LR", "SR"):
            assert torch.equal(torch.cat([p[i][0][k] for p in parts]), full[i][0][k]), (i, k)
        assert sum((p[i][1] for p in parts), []) == full[i][1]
    with pytest.raises(AssertionError, match="divide"):
        L(train_set, 6, device=None, shard=(0, 4))


def test_loader_abandoned_iteration_releases_its_reader_thread(handler):
    """Breaking out of an epoch early (the reference's ``next(iter(val_loader))`` in sample.py:79) must not leave the staging thread
    blocked: closing the generator stops it."""
    import threading
    import time
    L = wsr.sub("data.dataset_builder").DeviceBatchLoader
    train_set, _ = handler.get_datasets()
    before = threading.active_count()
    it = iter(L(train_set, 8, shuffle=True, num_workers=2, device=None, seed=1))
    next(it)
    next(it)
    it.close()
    deadline = time.time() + 10
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() <= before
    # a reader failure surfaces in the consumer instead of hanging it
    broken = L(train_set, 8, device=None)
    broken._stage = lambda *a, **k: (_ for _ in ()).throw(OSError("disk gone"))
    with pytest.raises(OSError, match="disk gone"):
        next(iter(broken))
