"""Measurement of the steps either side of the sampling loop on one B200.

    python tools/prof_edge.py [--hours 768] [--batch 64] [--workers 8]

1. the three HBM-bound edge kernels at the benchmark shape (B = 64, 128x256): achieved GB/s over algorithmic bytes
   (bicubic: read LR + write SR; standard scale: read + write; error sums: two reads), CUDA-event timed, L2 flushed between
   iterations by cycling through buffers whose total exceeds the 126 MB L2;
2. the store loader: batches/s of ``DeviceBatchLoader`` (pinned staging + side-stream copy + three launches) against the
   reference's approach on the same store and the same dataset objects -- ``torch.utils.data.DataLoader`` with worker
   processes over the per-sample interface, per-sample transform, ``torch.cat`` collate with CPU bicubic, then ``.cuda()``.
   The store is synthetic (random fields at the WeatherBench t2m grid sizes 32x64 / 128x256) written to a temporary directory."""
import argparse
import json
import os
import sys
import tempfile
import time
from datetime import datetime, timedelta

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wsr

nat = wsr.pkg.native


def write_store(root, hours, lr=(32, 64), scale=4):
    rng = np.random.default_rng(0)
    t0 = datetime(2000, 1, 1)
    for kind, (h, w) in (("lr", lr), ("hr", (lr[0] * scale, lr[1] * scale))):
        base = os.path.join(root, kind, "t2m")
        os.makedirs(os.path.join(base, "meta"))
        with open(os.path.join(base, "meta", "metadata.json"), "w") as fh:
            json.dump({"name": "t2m", "time_variate": True, "dims": ["lat", "lon"], "shape": [h, w],
                       "coords": [{"name": "lat", "values": list(range(h)), "dims": ["lat"]},
                                  {"name": "lon", "values": list(range(w)), "dims": ["lon"]}], "attrs": {}}, fh)
        for i in range(hours):
            t = t0 + timedelta(hours=i)
            d = os.path.join(base, "samples", str(t.year))
            os.makedirs(d, exist_ok=True)
            np.save(os.path.join(d, t.strftime("%Y-%m-%d-%H") + ".npy"), (280 + 10 * rng.standard_normal((h, w))).astype(np.float32))


def timed(fn, sets, reps=20):
    """fn(i) is run on buffer set i % len(sets) so consecutive iterations never find their inputs in L2."""
    for i in range(3):
        fn(i % sets)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        fn(i % sets)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def kernels(B):
    builder, transforms, metrics = wsr.sub("data.dataset_builder"), wsr.sub("data.transforms"), wsr.sub("training.metrics")
    dev = torch.device("cuda:0")
    sets = 6 if B <= 64 else 3
    lr = [torch.randn(B, 1, 32, 64, device=dev) for _ in range(sets)]
    hr = [torch.randn(B, 1, 128, 256, device=dev) for _ in range(sets)]
    hr2 = [torch.randn(B, 1, 128, 256, device=dev) for _ in range(sets)]
    mean, std = torch.full((B, 1), 280.0), torch.full((B, 1), 9.0)
    mean, std = mean.to(dev), std.to(dev)
    out = []
    nb_hr = B * 128 * 256 * 4
    ms = timed(lambda i: builder.bicubic_sr(lr[i], 4), sets)
    out.append(("wsr_bicubic_upsample  (B=%d, 32x64 -> 128x256)" % B, ms, (nb_hr + nb_hr // 16) / ms / 1e6))
    ms = timed(lambda i: transforms.inverse_batch(hr[i], mean, std), sets)
    out.append(("wsr_standard_scale    (B=%d, 128x256, per-sample statistics)" % B, ms, 2 * nb_hr / ms / 1e6))
    sums = metrics.ErrorSums(dev)
    ms = timed(lambda i: sums.update(hr[i], hr2[i], scale=std.reshape(-1)), sets)
    out.append(("wsr_error_sums        (B=%d, 128x256, inverse transform folded in)" % B, ms, 2 * nb_hr / ms / 1e6))
    return out


def loaders(root, hours, batch, workers):
    builder, transforms = wsr.sub("data.dataset_builder"), wsr.sub("data.transforms")
    t0 = datetime(2000, 1, 1)
    end = (t0 + timedelta(hours=hours)).strftime("%Y-%m-%d-%H")
    months = sorted({(t0 + timedelta(hours=i)).month for i in range(hours)})
    dh = builder.DataHandler(root, ["t2m"], root, months, [months], transforms.GlobalStandardScaling, "2000-01-01-00", end,
                             "2000-01-01-00", end, batch, batch, True, workers)
    t_fit = time.perf_counter()
    dh.process_data()
    t_fit = time.perf_counter() - t_fit
    train_set = dh.train_dataset
    res = {"fit_and_index_seconds": t_fit, "samples": len(train_set)}

    def run(loader, to_device):
        torch.cuda.synchronize()
        t = time.perf_counter()
        n = 0
        for b, _ in loader:
            if to_device:
                b = {k: v.cuda(non_blocking=True) for k, v in b.items()}
            n += b["HR"].shape[0]
        torch.cuda.synchronize()
        return n / (time.perf_counter() - t)

    dev_loader = builder.DeviceBatchLoader(train_set, batch, shuffle=True, num_workers=workers, device="cuda:0")
    run(dev_loader, False)                                    # page cache warm for both arms
    res["device_loader_samples_per_s"] = run(dev_loader, False)
    ref_style = torch.utils.data.DataLoader(train_set, batch_size=batch, collate_fn=lambda s: builder.form_batch(s, 4, None), shuffle=True,
                                            pin_memory=True, drop_last=True, num_workers=workers)
    res["reference_style_dataloader_samples_per_s"] = run(ref_style, True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=int, default=768)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--workers", type=int, default=8)
    a = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6512.0)
    except Exception:
        peak = 6512.0
    print("# edge kernels (CUDA events, inputs cycled through %d buffer sets > L2); HBM peak %.0f GB/s" % (6, peak))
    for bsz in (a.batch, 1024):          # the benchmark batch (8 MB tensors: launch-bound) and a device-bound size (134 MB tensors)
        for name, ms, gbs in kernels(bsz):
            print("%-70s %.4f ms  %7.0f GB/s  %.0f%% of peak" % (name, ms, gbs, 100 * gbs / peak))
        torch.cuda.empty_cache()
    with tempfile.TemporaryDirectory() as root:
        write_store(root, a.hours)
        r = loaders(root, a.hours, a.batch, a.workers)
    print("# store loader, %d samples (32x64 LR + 128x256 HR fp32), batch %d, %d reader threads / worker processes, %d host cores" % (
        r["samples"], a.batch, a.workers, os.cpu_count()))
    print("index + fit (GlobalStandardScaling, lr + hr)           %.2f s" % r["fit_and_index_seconds"])
    print("DeviceBatchLoader (device-resident HR/LR/SR)            %.0f samples/s" % r["device_loader_samples_per_s"])
    print("reference-style DataLoader + CPU bicubic + .cuda()      %.0f samples/s" % r["reference_style_dataloader_samples_per_s"])
    print("denoiser consumption at the benchmark rate for comparison: training 158 samples/s/GPU, sampling 3.6 samples/s/GPU")


if __name__ == "__main__":
    main()
