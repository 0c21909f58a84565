#!/usr/bin/env python
"""bench.py -- headline benchmark of the diffusion super-resolution hot path (BASELINE.json).

Metric: 128x256 t2m super-resolution samples/s for the full 1000-step DDPM reverse loop (ResDiff Cfg-A UNet, bf16,
batch 64 per GPU, synthetic WeatherBench-shaped fields, random-init weights).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores
    python bench.py --impl reference --ref-device cuda [--ref-batch 16] [--ref-autocast]     # optional extra: the same algorithm run
                                                             # op by op by eager PyTorch on cuda:0 (never the default arm)
    python bench.py --workload train [--profile-ops]         # configs[2]: the training step

A "step" is ONE reverse (p_sample) step of the loop over the whole local batch: level-projection select, FD gate,
stem assembly, the UNet (~270 kernel launches, replayed as one CUDA graph), the fused sampler update.  Every one of the
1000 steps of a chain is the same work, so
    value = samples/s = global_batch / (T * seconds_per_step + once-per-batch precompute seconds),  T = 1000,
with K timed steps (K = 1000 times the whole chain).  The activations of one step (GBs at batch 64) far exceed the
126 MB L2, so no L2 flush is needed between timed steps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_FULL = 1000
FLOPS_PER_SAMPLE_STEP = 206.39e9        # algorithmic FLOPs of one ResDiff Cfg-A UNet call per sample
LINEAR_1000 = {"schedule": "linear", "n_timestep": T_FULL, "linear_start": 1e-6, "linear_end": 1e-2}
CFG_A = dict(in_channel=5, out_channel=1, norm_groups=32, inner_channel=64, channel_mults=[1, 2, 4, 8, 8], attn_res=[16],
             res_blocks=2, dropout=0.2, image_height=128, image_width=256, image_channels=1)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on stdout when NCCL_DEBUG
    is set, as it is on the GPU boxes): file descriptor 1 is pointed at stderr for the rest of the run and the JSON line goes to the
    saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(text):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(n_steps, warmup, seed=0):
    """Times reverse steps of the reference algorithm (oracle/ port, fp32 torch-CPU, all host threads) at B=1, 128x256.
    Returns (seconds per step list, cores)."""
    import torch
    from oracle import nets, process
    from oracle.schedule import ddpm_tables
    from oracle.weights import seeded_randn, seeded_state_dict
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    man = np.load(os.path.join(ROOT, "tests", "golden", "manifest.npz"))
    keys = [str(k) for k in man["resdiff.keys"]]
    shapes = [tuple(int(x) for x in str(s).split(",")) if str(s) else () for s in man["resdiff.shapes"]]
    sd = seeded_state_dict(zip(keys, shapes), seed)
    cfg = dict(CFG_A)
    cfg["dropout"] = 0.0
    tab32, sap = ddpm_tables(LINEAR_1000)
    tab = {k: torch.from_numpy(v) for k, v in tab32.items()}
    cond = torch.nn.functional.interpolate(seeded_randn("bench.lr", (1, 1, 32, 64), 1234), scale_factor=4, mode="bicubic")
    x = seeded_randn("bench.x", (1, 1, 128, 256), 4321)

    def denoise(xx, level):
        return nets.resdiff_unet(sd, torch.cat([cond, xx], 1), level, cfg)

    times = []
    t = T_FULL - 1
    with torch.no_grad():
        for i in range(warmup + n_steps):
            z = seeded_randn("bench.z%d" % i, (1, 1, 128, 256), 99)
            t0 = time.perf_counter()
            x, _ = process.p_sample_step(denoise, tab, sap, x, t, z)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            t -= 1
    return times, cores


def gpu_eager_reference_steps(n_steps, warmup, batch, autocast, seed=0):
    """OPTIONAL extra arm (``--impl reference --ref-device cuda``, never the default): the same oracle port of the reference's
    algorithm, but executed by eager PyTorch (cuDNN / cuBLAS, one launch per op, fp32 or bf16 autocast) on cuda:0 at ``batch``
    samples -- what running the reference's own modules on a B200 amounts to.  Returns seconds per step (CUDA events)."""
    import torch
    from oracle import nets, process
    from oracle.schedule import ddpm_tables
    from oracle.weights import seeded_randn, seeded_state_dict
    import numpy as np
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True                    # the reference sets it (train.py:24-25)
    man = np.load(os.path.join(ROOT, "tests", "golden", "manifest.npz"))
    keys = [str(k) for k in man["resdiff.keys"]]
    shapes = [tuple(int(x) for x in str(s).split(",")) if str(s) else () for s in man["resdiff.shapes"]]
    sd = {k: v.to(dev) for k, v in seeded_state_dict(zip(keys, shapes), seed).items()}
    cfg = dict(CFG_A)
    cfg["dropout"] = 0.0
    tab32, sap = ddpm_tables(LINEAR_1000)
    tab = {k: torch.from_numpy(v).to(dev) for k, v in tab32.items()}
    cond = torch.nn.functional.interpolate(seeded_randn("bench.lr", (batch, 1, 32, 64), 1234), scale_factor=4, mode="bicubic").to(dev)
    x = seeded_randn("bench.x", (batch, 1, 128, 256), 4321).to(dev)
    z = seeded_randn("bench.z", (batch, 1, 128, 256), 99).to(dev)

    def denoise(xx, level):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return nets.resdiff_unet(sd, torch.cat([cond, xx], 1), level.to(dev), cfg).float()

    times = []
    t = T_FULL - 1
    with torch.no_grad():
        for i in range(warmup + n_steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x, _ = process.p_sample_step(denoise, tab, sap, x, t, z)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                times.append(e0.elapsed_time(e1) * 1e-3)
            t -= 1
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        times = gpu_eager_reference_steps(args.steps, max(args.warmup, 2), args.ref_batch, args.ref_autocast)
        ms = 1e3 * sum(times) / len(times)
        val = args.ref_batch / (T_FULL * ms * 1e-3)
        _emit(json.dumps({
            "impl": "reference", "device": "cuda:0 eager PyTorch (%s)" % ("bf16 autocast" if args.ref_autocast else "fp32"),
            "metric": "128x256 t2m SR samples/sec (1000-step loop)", "value": val, "unit": "samples/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.ref_autocast else "f32", "data": "synthetic",
            "config": {"workload": "ResDiff Cfg-A UNet, 128x256, 1000-step DDPM reverse loop; oracle port of the reference run by eager PyTorch",
                       "batch_per_gpu": args.ref_batch, "T": T_FULL},
            "note": "extra arm, not the CPU baseline: same algorithm, library kernels (cuDNN / cuBLAS) launched op by op", "gpu_launches": 0}))
        return
    times, cores = cpu_reference_steps(args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    val = 1.0 / (T_FULL * ms * 1e-3)
    sample = "B=1 of the batch, %d reverse steps of the 1000-step chain at 128x256 (oracle port of the reference, fp32 torch-CPU)" % len(times)
    line = {
        "impl": "reference", "metric": "128x256 t2m SR samples/sec (1000-step loop)", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResDiff Cfg-A UNet, 128x256, 1000-step DDPM reverse loop; CPU arm runs batch 1", "T": T_FULL},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def _conv_traffic():
    """DRAM bytes per gemm_tc_kernel launch (read + write, averaged over the launches of a B=64 sampler step) from the
    committed ncu pass (profiles/r01c_conv_traffic.json <- profiles/r01c_ncu_launches.csv); None if the file is missing."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01c_conv_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import wsr
    nat = wsr.pkg.native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    B = args.batch if args.scaling == "weak" else max(1, args.batch // world)
    global_batch = B * world

    U = wsr.sub("models.diffusion_models.resdiff.unet").UNet
    D = wsr.sub("models.diffusion_models.resdiff.resdiff_diffusion").ResDiffDiffusion
    networks = wsr.sub("models.diffusion_models.networks")
    torch.manual_seed(0)
    net = U(precision="bf16", **CFG_A)
    networks.init_weights(net, "orthogonal")          # the reference's "random-init weights" (networks.py:164-165)
    net = net.to(dev).eval()
    diff = D(net, image_height=128, image_width=256, channels=1, conditional=True).to(dev)
    diff.set_new_noise_schedule(LINEAR_1000, dev)

    # synthetic WeatherBench-shaped standardised fields: LR randn, SR = bicubic x4, in PINNED host memory
    g = torch.Generator().manual_seed(1234 + rank)
    lr = torch.randn(B, 1, 32, 64, generator=g)
    sr_host = torch.nn.functional.interpolate(lr, scale_factor=4, mode="bicubic").contiguous().pin_memory()
    out_host = torch.empty_like(sr_host).pin_memory()

    plan = net.plan(B, dev, strict_tc=False)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- once-per-batch precompute (condition-only hoists + level table + initial noise) -------------------------------
    cond = sr_host.to(dev, non_blocking=True)
    torch.cuda.synchronize(dev)
    e0, e1 = ev(), ev()
    plan.set_condition(cond)                       # untimed first call (lazy init)
    loop = diff.begin_loop(plan, tuple(cond.shape), seed=7)
    torch.cuda.synchronize(dev)
    e0.record()
    plan.set_condition(cond)
    loop = diff.begin_loop(plan, tuple(cond.shape), seed=7)
    e1.record()
    torch.cuda.synchronize(dev)
    precompute_ms = e0.elapsed_time(e1)

    # ---- warm-up (eager, also counts launches per step), capture, timed graph replays ----------------------------------
    l0 = nat.launches
    loop.step()
    launches_per_step = nat.launches - l0
    for _ in range(max(0, args.warmup - 1)):
        loop.step()
    loop.capture()
    loop.replay()                                   # one graph warm-up replay
    K = args.steps
    if K + args.warmup + 2 > T_FULL:
        K = T_FULL - args.warmup - 2
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(K):
        loop.replay()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_local = e0.elapsed_time(e1)
    tmax = torch.tensor([ms_local, precompute_ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, pre_ms = float(tmax[0]), float(tmax[1])
    ms_per_step = ms_total / K
    value = global_batch / (T_FULL * ms_per_step * 1e-3 + pre_ms * 1e-3)

    # ---- roofline pass: one eager step with per-launch CUDA events ------------------------------------------------------
    plan.eng.prof = []
    loop.step()
    summ = plan.eng.prof_summary()
    plan.eng.prof = None
    peaks, peak_src = _peaks()
    conv = summ.get("conv_tc", [0, 1e-9, 0, 0])
    conv_tflops = conv[2] / (conv[1] * 1e-3) / 1e12 if conv[1] > 0 else 0.0
    step_ms_prof = sum(v[1] for v in summ.values())
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc_kernel (conv_tc launches)", "achieved": conv_tflops, "peak": peaks["bf16_tflops"],
        "unit": "TFLOP/s", "frac": conv_tflops / peaks["bf16_tflops"], "traffic": _conv_traffic(),
        "traffic_note": "ncu dram__bytes_read+write per gemm_tc_kernel launch, averaged over one B=64 step (profiles/r01c_conv_traffic.json)",
        "peak_source": peak_src + " burst",
        "launches_per_step": conv[0], "kernel_ms_per_step": conv[1], "kernel_share_of_step": conv[1] / step_ms_prof if step_ms_prof else None,
        "whole_step_tflops": global_batch / world * FLOPS_PER_SAMPLE_STEP / (ms_per_step * 1e-3) / 1e12,
        "whole_step_frac_of_sustained": global_batch / world * FLOPS_PER_SAMPLE_STEP / (ms_per_step * 1e-3) / 1e12 / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
    }
    if args.profile_ops and rank == 0:
        plan.eng.prof, plan.eng.prof_detail = [], True
        loop.step()
        detail = plan.eng.prof_summary()
        plan.eng.prof, plan.eng.prof_detail = None, False
        print("# per-shape breakdown of the tensor-core launches (B=%d): name launches ms TFLOP/s" % B, file=sys.stderr)
        for name, (n, ms, fl, nb) in sorted(detail.items(), key=lambda kv: -kv[1][1]):
            if name.startswith("conv_tc") or name.startswith("gemm_tc"):
                print("#   %-40s %3d %8.3f %8.1f" % (name, n, ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0), file=sys.stderr)
        rows = sorted(summ.items(), key=lambda kv: -kv[1][1])
        print("# per-op breakdown of one eager step (B=%d): name launches ms TFLOP/s GB/s" % B, file=sys.stderr)
        for name, (n, ms, fl, nb) in rows:
            print("#   %-22s %4d %9.3f %9.1f %9.1f" % (name, n, ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0, nb / (ms * 1e-3) / 1e9 if ms > 0 else 0), file=sys.stderr)
        print("#   total %.3f ms (graph: %.3f ms/step)" % (step_ms_prof, ms_per_step), file=sys.stderr)

    # ---- end-to-end through the public API with HOST buffers ------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        budget_s = args.e2e_budget
        t_e2e = T_FULL
        est = T_FULL * ms_per_step * 1e-3
        if est > budget_s:
            t_e2e = max(10, int(T_FULL * budget_s / est))
        sched = dict(LINEAR_1000)
        sched["n_timestep"] = t_e2e
        diff.set_new_noise_schedule(sched, dev)
        diff.sample_seed = 11
        barrier()
        t0 = time.perf_counter()
        x_in = {"SR": sr_host.to(dev, non_blocking=True)}
        res = diff.super_resolution(x_in)
        out_host.copy_(res, non_blocking=True)
        torch.cuda.synchronize(dev)
        el = time.perf_counter() - t0
        tt = torch.tensor([el], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el = float(tt[0])
        e2e = {"value": global_batch / (el * T_FULL / t_e2e), "unit": "samples/s", "h2d_bytes_per_step": sr_host.numel() * 4,
               "d2h_bytes_per_step": out_host.numel() * 4, "api": "ResDiffDiffusion.super_resolution({'SR': host tensor}) + .cpu()",
               "loop_steps_run": t_e2e, "seconds": el,
               "note": "one public-API call = one whole reverse loop; H2D of the condition and D2H of the result inside the timed region"
                       + ("" if t_e2e == T_FULL else "; run with a %d-step schedule and scaled to 1000 steps" % t_e2e)}
        diff.set_new_noise_schedule(LINEAR_1000, dev)

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        times, cores = cpu_reference_steps(args.cpu_steps, 1)
        mean = sum(times) / len(times)
        cpu = {"value": 1.0 / (T_FULL * mean), "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": "B=1, %d reverse steps at 128x256 of the oracle port (fp32 torch-CPU), %.2f s/step" % (len(times), mean)}

    if rank == 0:
        line = {
            "metric": "128x256 t2m SR samples/sec (1000-step loop)", "value": value, "unit": "samples/s", "n_gpus": n_gpus,
            "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: ResDiff Cfg-A UNet (inner 64, mults 1-2-4-8-8, attn@16) + bicubic prior, t2m 32x64->128x256, "
                                   "1000-step DDPM reverse loop, bf16", "global_batch": global_batch, "batch_per_gpu": B, "T": T_FULL,
                       "step": "one reverse step over the local batch (CUDA-graph replay)", "precompute_ms_per_batch": pre_ms,
                       "l2": "per-step activations >> 126 MB L2, no flush needed", "parallelism": "batch-sharded x%d, no collective in the loop" % world},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu,
            "tc_launches_per_step": plan.eng.n_tc, "simt_launches_total": plan.eng.n_simt,
        }
        _emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# secondary workload: configs[2], the training step (not the headline metric; `--workload train`)
# ----------------------------------------------------------------------------------------------------------------------
def run_train(args):
    """One training step = q_sample + UNet forward (dropout 0.2) + loss + hand-written backward + bucketed gradient
    all-reduce + one-launch Adam over the flat parameter buffer, batch 4 per GPU at 128x256 (BASELINE.json configs[2])."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import wsr
    nat = wsr.pkg.native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.train_batch
    U = wsr.sub("models.diffusion_models.resdiff.unet").UNet
    D = wsr.sub("models.diffusion_models.resdiff.resdiff_diffusion").ResDiffDiffusion
    glue = wsr.sub("autograd_glue")
    par = wsr.sub("parallel")
    networks = wsr.sub("models.diffusion_models.networks")
    torch.manual_seed(0); np.random.seed(rank)
    net = U(precision=args.train_precision, **CFG_A)
    networks.init_weights(net, "orthogonal")
    net = net.to(dev).train()
    diff = D(net, image_height=128, image_width=256, channels=1, conditional=True).to(dev)
    diff.set_new_noise_schedule(LINEAR_1000, dev)
    diff.set_loss(dev)
    plan = net.train_plan(B, dev)
    opt = glue.FusedAdam(list(diff.parameters()), lr=1e-4)
    opt.attach_flat(plan)
    reducer = par.FlatGradReducer(plan) if world > 1 else None
    g = torch.Generator().manual_seed(1234 + rank)
    lr = torch.randn(B, 1, 32, 64, generator=g)
    sr = torch.nn.functional.interpolate(lr, scale_factor=4, mode="bicubic").to(dev)
    hr = sr + 0.3 * torch.randn(sr.shape, generator=g).to(dev)
    numel = hr.numel() * world

    def step():
        opt.zero_grad()
        loss = diff.p_losses({"HR": hr, "SR": sr})
        (loss.sum() / numel).backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = nat.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = nat.launches - l0
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    if args.profile_ops and rank == 0:
        plan.eng.prof = []
        step()
        summ = plan.eng.prof_summary()
        plan.eng.prof = None
        tot = sum(v[1] for v in summ.values())
        print("# per-op breakdown of one training step (B=%d, %s): name launches ms TFLOP/s" % (B, args.train_precision), file=sys.stderr)
        for name, (n, t, fl, nb) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
            print("#   %-22s %4d %9.3f %9.1f" % (name, n, t, fl / (t * 1e-3) / 1e12 if t > 0 else 0), file=sys.stderr)
        print("#   total of timed kernels %.3f ms (wall per step %.3f ms)" % (tot, ms), file=sys.stderr)
        # per-shape detail: the recorded launch lists carry no shape tags, so this one step is issued call by call
        plan.replay_enabled = False
        plan.eng.prof, plan.eng.prof_detail = [], True
        step()
        detail = plan.eng.prof_summary()
        plan.eng.prof, plan.eng.prof_detail = None, False
        plan.replay_enabled = True
        print("# per-shape breakdown of the tensor-core launches of one training step: name launches ms TFLOP/s", file=sys.stderr)
        for name, (n, t, fl, nb) in sorted(detail.items(), key=lambda kv: -kv[1][1])[:90]:
            if name.startswith(("wgrad_tc", "conv_tc", "gemm_tc", "conv_taps")):
                print("#   %-42s %3d %8.3f %8.1f" % (name, n, t, fl / (t * 1e-3) / 1e12 if t > 0 else 0), file=sys.stderr)
            elif name.startswith("gn_bwd"):
                print("#   %-42s %3d %8.3f %8.1f GB/s" % (name, n, t, nb / (t * 1e-3) / 1e9 if t > 0 else 0), file=sys.stderr)
    if rank == 0:
        peaks, peak_src = _peaks()
        tfl = 3 * FLOPS_PER_SAMPLE_STEP * B / (ms * 1e-3) / 1e12
        _emit(json.dumps({
            "metric": "ResDiff training step samples/sec (fwd+bwd+allreduce+Adam)", "value": B * world / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.train_precision, "data": "synthetic",
            "config": {"workload": "configs[2]: ResDiff Cfg-A training step, 128x256, dropout 0.2, batch %d per GPU" % B,
                       "global_batch": B * world, "parallelism": "data-parallel x%d, bucketed gradient all-reduce overlapped with backward" % world},
            "clocks": clocks, "gpu_launches": launches, "loss": float(loss),
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tfl / peaks["bf16_tflops"],
                         "note": "3 x forward algorithmic FLOPs (SURVEY 8d) / step time, per GPU", "peak_source": peak_src}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="reference arm: cpu (the contract) or the optional eager-PyTorch-on-GPU arm")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--ref-autocast", action="store_true")
    ap.add_argument("--batch", type=int, default=64, help="batch per GPU (weak) or total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-budget", type=float, default=60.0, help="seconds allowed for the end-to-end public-API call")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--profile-ops", action="store_true")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"], help="sample = headline metric; train = configs[2]")
    ap.add_argument("--train-batch", type=int, default=4)
    ap.add_argument("--train-precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
