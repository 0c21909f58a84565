"""SURVEY 8f N3 on the device: ``DeviceBatchLoader`` (raw fields read straight into pinned memory, one H2D copy per group on a
side stream, batch standardisation and the bicubic condition as three launches) against the fixture produced by the REAL
reference DataHandler (tests/golden/store.npz) and against the host-side loader over whole epochs."""
import pytest
import torch

import wsr
from conftest import load_golden
from oracle import store

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handler(tmp_path_factory):
    root = store.write_store(str(tmp_path_factory.mktemp("store")))
    builder, transforms = wsr.sub("data.dataset_builder"), wsr.sub("data.transforms")
    s = store.SPEC
    dh = builder.DataHandler(root, list(store.VARIABLES), root, s["months_subset"], s["groups"], transforms.GlobalStandardScaling,
                             s["train"][0], s["train"][1], s["val"][0], s["val"][1], s["val_batch_size"], s["train_batch_size"], False, 4)
    dh.process_data()
    return dh


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def test_device_loader_matches_reference_fixture(handler):
    g = load_golden("store")
    assert handler.device is not None and handler.device.type == "cuda"
    launches0 = wsr.pkg.native.launches
    batch, months = next(iter(handler.val_loader))
    assert wsr.pkg.native.launches - launches0 == 3          # standardise LR, standardise HR, bicubic condition
    assert months == [int(m) for m in g["val0.months"]]
    for k in ("HR", "LR", "SR"):
        assert batch[k].is_cuda and batch[k].shape == g["val0." + k].shape
        assert _rel(batch[k], g["val0." + k]) < 2e-5, k
    by_date, bm = handler.get_data_by_date("2000-03-05-07")
    for k in ("HR", "LR", "SR"):
        assert by_date[k].is_cuda and _rel(by_date[k], g["date." + k]) < 2e-5, k
    inv = handler.get_data_transformer().inverse_transform(by_date, bm)
    for k in ("HR", "LR", "SR"):
        assert inv[k].is_cuda and _rel(inv[k], g["date_inv." + k]) < 1e-6, k


def test_device_epochs_equal_host_epochs(handler):
    L = wsr.sub("data.dataset_builder").DeviceBatchLoader
    train_set, _ = handler.get_datasets()
    for shuffle in (False, True):
        dev = L(train_set, 32, shuffle=shuffle, num_workers=4, device="cuda:0", seed=11, prefetch=2)
        host = L(train_set, 32, shuffle=shuffle, num_workers=1, device=None, seed=11)
        n = 0
        for (bd, md), (bh, mh) in zip(dev, host):
            assert md == mh
            for k in ("HR", "LR", "SR"):
                assert _rel(bd[k], bh[k]) < 2e-6, (k, n)
            n += 1
        assert n == len(dev) == 816 // 32


def test_loader_feeds_the_model(handler):
    """One validation batch from the store through the reference-facing model classes: sample, then RMSE in Kelvin with the
    fitted statistics folded into the metric pass."""
    from oracle.cases import unet_cfg
    cfg = unet_cfg(32, 64, inner=64, c_img=2, attn_res=(4,))
    U = wsr.sub("models.diffusion_models.sr3.unet").UNet
    D = wsr.sub("models.diffusion_models.sr3.sr3_diffusion").SR3Diffusion
    metrics = wsr.sub("training.metrics")
    torch.manual_seed(0)
    net = U(in_channel=4, out_channel=2, norm_groups=32, inner_channel=64, channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
            res_blocks=2, dropout=0, image_height=32, image_width=64, image_channels=2).cuda().eval()
    diff = D(net, image_height=32, image_width=64, channels=2, conditional=True).cuda()
    diff.set_new_noise_schedule({"schedule": "linear", "n_timestep": 4, "linear_start": 1e-4, "linear_end": 2e-2}, torch.device("cuda:0"))
    batch, months = next(iter(handler.val_loader))
    sr = diff.super_resolution(batch, False)
    assert sr.shape == batch["HR"].shape and bool(torch.isfinite(sr).all())
    _, std = handler.get_data_transformer().batch_statistics("hr", months)
    vm = metrics.ValidationMetrics(metrics.create_metric_dict("cuda:0"))
    vm.update(sr, batch["HR"], scale=std.reshape(-1))
    vm.compute_metrics()
    inv = handler.get_data_transformer().inverse_transform({"SR": sr, "HR": batch["HR"]}, months)
    rmse_k = float(((inv["SR"] - inv["HR"]) ** 2).mean().sqrt())
    assert float(vm.metrics2dict()["RMSE"]) == pytest.approx(rmse_k, rel=1e-4)
