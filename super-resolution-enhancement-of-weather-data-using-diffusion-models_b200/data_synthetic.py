"""Batch source for the entry points.  The reference's data pipeline (WeatherBench .npy store reader, transforms,
DataLoader collate: data/, 2245 lines) is host-side I/O outside the accelerated path; what the model
consumes is only the batch-dict contract of data/dataset_builder.py:344-382:

    {'HR': (B,C,H,W), 'LR': (B,C,H/4,W/4), 'SR': bicubic x4 of LR (B,C,H,W)} fp32 standardised units, + months list

This module produces that contract from the reference's on-disk store when ``data.dataroot`` is one (``<root>/{lr,hr}/<var>/...``,
SURVEY 8f N3: ``data.dataset_builder.DataHandler`` + the device loader), else from synthetic WeatherBench-shaped fields or from a
file of real, already standardised fields (``.npz`` / ``.pt`` with arrays ``HR`` and ``LR``)."""
import os

import numpy as np
import torch
import torch.nn.functional as F


def form_batch(lr, hr=None, scale=4):
    if lr.is_cuda:
        from .data.dataset_builder import bicubic_sr
        sr = bicubic_sr(lr, scale)                                   # device-side collate
    else:
        sr = F.interpolate(lr, scale_factor=scale, mode="bicubic")   # dataset_builder.py:377 (host data source)
    if hr is None:
        hr = sr.clone()
    return {"HR": hr, "LR": lr, "SR": sr}, [1] * lr.shape[0]


def synthetic_batches(n_batches, batch, channels, height, width, scale=4, seed=1234):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n_batches):
        lr = torch.randn(batch, channels, height // scale, width // scale, generator=g)
        sr = F.interpolate(lr, scale_factor=scale, mode="bicubic")
        hr = sr + 0.3 * torch.randn(batch, channels, height, width, generator=g)
        yield form_batch(lr, hr, scale)


def file_batches(path, batch):
    if path.endswith(".npz"):
        z = np.load(path)
        hr, lr = torch.from_numpy(z["HR"]).float(), torch.from_numpy(z["LR"]).float()
    else:
        z = torch.load(path)
        hr, lr = z["HR"].float(), z["LR"].float()
    for i in range(0, hr.shape[0], batch):
        yield form_batch(lr[i:i + batch], hr[i:i + batch], hr.shape[-1] // lr.shape[-1])


_handlers = {}


def is_store(root):
    return bool(root) and os.path.isdir(os.path.join(root, "lr")) and os.path.isdir(os.path.join(root, "hr"))


def store_handler(opt, val_only=False, shard=None):
    """The reference's ``DataHandler(...)`` call of train.py:232-238 / sample.py:51-56 for ``opt['data']``; None when
    ``dataroot`` is not a store directory.  ``val_only`` = sample.py's variant (transforms fitted on the validation range)."""
    d = opt["data"]
    root = str(d.get("dataroot", ""))
    if not is_store(root):
        return None
    key = (root, val_only, shard, str(d.get("train_min_date")), str(d.get("train_max_date")), str(d.get("val_min_date")), str(d.get("val_max_date")),
           str(d.get("months_subset")))
    if key not in _handlers:
        from .data.dataset_builder import DataHandler
        from .data.transforms import get_transformation_by_name
        groups = d["transform_groups"]
        groups = list(groups.values()) if isinstance(groups, dict) else groups
        lo, hi = (d["val_min_date"], d["val_max_date"]) if val_only else (d["train_min_date"], d["train_max_date"])
        storage = opt.get("path", {}).get("experiments_root") or root
        if not os.path.isdir(storage) or not os.access(storage, os.W_OK):
            import tempfile
            storage = tempfile.mkdtemp(prefix="wsr_meta_")
        dh = DataHandler(root, d["variables"], storage, d["months_subset"], groups, get_transformation_by_name(d["transformation"]),
                         lo, hi, d["val_min_date"], d["val_max_date"], d["val_batch_size"], d["batch_size"], d.get("use_shuffle", True),
                         d.get("num_workers", 4), shard=shard)
        dh.process_data()
        _handlers[key] = dh
    return _handlers[key]


def batches_from_opt(opt, phase, n_batches=None, shard=None):
    """shard = (rank, world): the store's training loader then yields this rank's slice of every global batch (already sharded);
    synthetic / file batches are always global and are cut by the caller."""
    d = opt["data"]
    m = opt["model"]["diffusion"]
    bs = d["val_batch_size"] if phase == "val" else d["batch_size"]
    root = str(d.get("dataroot", ""))
    dh = store_handler(opt, shard=shard)
    if dh is not None:
        if phase == "val":
            return iter(dh.val_loader)

        def epochs():                          # train.py:57-64: epochs over the loader until n_iter batches were served
            served, limit = 0, (n_batches if n_batches is not None else opt["train"]["n_iter"])
            while served < limit:
                for item in dh.train_loader:
                    if served >= limit:
                        return
                    served += 1
                    yield item
        return epochs()
    if root.endswith((".npz", ".pt")):
        return file_batches(root, bs)
    n = n_batches if n_batches is not None else (1 if phase == "val" else opt["train"]["n_iter"])
    return synthetic_batches(n, bs, m["image_channels"], m["image_height"], m["image_width"])


def epoch_batches(opt, remaining, shard=None):
    """ONE epoch of training batches, at most ``remaining`` of them (reference train.py:57-64: ``for train_data in train_loader``
    inside ``while curr_iter <= n_iter``).  Store: one pass over the shuffled loader; file: one pass over the file; synthetic
    fields have no epoch structure, so all remaining batches form one epoch."""
    d = opt["data"]
    m = opt["model"]["diffusion"]
    root = str(d.get("dataroot", ""))
    dh = store_handler(opt, shard=shard)

    def limited(it):
        for i, item in enumerate(it):
            if i >= remaining:
                return
            yield item
    if dh is not None:
        return limited(dh.train_loader)
    if root.endswith((".npz", ".pt")):
        return limited(file_batches(root, d["batch_size"]))
    # the stream continues where a resumed run left off: the generator is advanced by seeding with the served count
    seed = 1234 + int(opt["train"]["n_iter"]) - int(remaining)
    return synthetic_batches(remaining, d["batch_size"], m["image_channels"], m["image_height"], m["image_width"], seed=seed)
