"""Profiling helper: runs one conv_tc shape a few times (for ncu / event timing).
usage: python tools/prof_conv.py N Cin Cout H W k [reps]"""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

nat = wsr.pkg.native
Engine = wsr.sub("engine").Engine


def main():
    N, Cin, Cout, H, W, k = [int(v) for v in sys.argv[1:7]]
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
    dev = torch.device("cuda:0")
    eng = Engine(dev, "bf16")
    torch.manual_seed(0)
    x = eng.new_act(N, H, W, Cin)
    x.buf.copy_(torch.randn_like(x.buf, dtype=torch.float32))
    w = torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, device=dev)
    pc = eng.pack_conv(w, b)
    arena = wsr.sub("engine").StatsArena()
    y = eng.new_act(N, H, W, Cout, stats=arena if os.environ.get("PROF_STATS") else None)
    arena.finalize(dev)
    rv = torch.randn(N, Cout, device=dev)
    res = x if (os.environ.get("PROF_RES") and Cin == Cout) else None
    for _ in range(2):
        eng.conv(x, pc, y, rowvec=rv.data_ptr(), rowvec_ld=Cout, res=res)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        eng.conv(x, pc, y, rowvec=rv.data_ptr(), rowvec_ld=Cout, res=res)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    fl = 2.0 * N * H * W * Cout * Cin * k * k
    print("conv %d->%d k%d %dx%d N=%d: %.3f ms  %.1f TFLOP/s  out+in %.1f GB/s" % (
        Cin, Cout, k, H, W, N, ms, fl / ms / 1e9, N * H * W * (Cin + Cout) * 2 / ms / 1e6))


if __name__ == "__main__":
    main()
