"""Batch source for the entry points.  The reference's data pipeline (WeatherBench .npy store reader, transforms,
DataLoader collate: data/, 2245 lines) is host-side I/O outside the accelerated path; what the model
consumes is only the batch-dict contract of data/dataset_builder.py:344-382:

    {'HR': (B,C,H,W), 'LR': (B,C,H/4,W/4), 'SR': bicubic x4 of LR (B,C,H,W)} fp32 standardised units, + months list

This module produces that contract either from synthetic WeatherBench-shaped fields or from a file of real,
already standardised fields (``.npz`` / ``.pt`` with arrays ``HR`` and ``LR``)."""
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


def batches_from_opt(opt, phase, n_batches=None):
    d = opt["data"]
    m = opt["model"]["diffusion"]
    bs = d["val_batch_size"] if phase == "val" else d["batch_size"]
    root = str(d.get("dataroot", ""))
    if root.endswith((".npz", ".pt")):
        return file_batches(root, bs)
    n = n_batches if n_batches is not None else (1 if phase == "val" else opt["train"]["n_iter"])
    return synthetic_batches(n, bs, m["image_channels"], m["image_height"], m["image_width"])
