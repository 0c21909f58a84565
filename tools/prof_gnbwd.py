"""GroupNorm backward kernels alone (for event timing / ncu).  usage: python tools/prof_gnbwd.py N H W C [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

nat = wsr.pkg.native
em = wsr.sub("engine")


def main():
    N, H, W, C = [int(v) for v in sys.argv[1:5]]
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    dev = torch.device("cuda:0")
    eng = em.Engine(dev, "bf16")
    arena = em.StatsArena()
    x = eng.new_act(N, H, W, C, stats=arena)
    arena.finalize(dev)
    x.buf.copy_(torch.randn_like(x.buf, dtype=torch.float32))
    eng.gn_stats(x)
    da, dx = eng.new_act(N, H, W, C), eng.new_act(N, H, W, C)
    da.buf.copy_(torch.randn_like(da.buf, dtype=torch.float32))
    dx.buf.zero_()
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    red = torch.zeros(N * 2 * C, dtype=torch.float64, device=dev)

    def run():
        red.zero_()
        eng.gn_bwd(x, g, b, 32, nat.ACT_SWISH, da, dx, red.data_ptr(), dg, db, True)
    run(); run()
    torch.cuda.synchronize()
    eng.prof = []
    for _ in range(reps):
        run()
    summ = eng.prof_summary()
    eng.prof = None
    for k, (n, ms, fl, nb) in summ.items():
        print("%-16s N=%d %dx%d C=%d: %.1f us per launch, %.0f GB/s" % (k, N, H, W, C, 1e3 * ms / n, nb / (ms * 1e-3) / 1e9))


if __name__ == "__main__":
    main()
