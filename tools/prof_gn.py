"""gn_apply bandwidth per shape.  usage: python tools/prof_gn.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

nat = wsr.pkg.native
em = wsr.sub("engine")


def main():
    dev = torch.device("cuda:0")
    eng = em.Engine(dev, "bf16")
    for (N, H, W, C) in [(64, 128, 256, 64), (64, 128, 256, 128), (64, 128, 256, 192), (64, 64, 128, 128), (64, 64, 128, 256), (64, 64, 128, 384),
                         (64, 32, 64, 256), (64, 32, 64, 768), (64, 16, 32, 512), (64, 16, 32, 1024), (64, 8, 16, 512), (64, 8, 16, 1024)]:
        arena = em.StatsArena()
        x = eng.new_act(N, H, W, C, stats=arena)
        arena.finalize(dev)
        x.buf.copy_(torch.randn_like(x.buf, dtype=torch.float32))
        eng.gn_stats(x)
        y = eng.new_act(N, H, W, C)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        for _ in range(2):
            eng.gn_apply(x, g, b, 32, nat.ACT_SWISH, y)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            eng.gn_apply(x, g, b, 32, nat.ACT_SWISH, y)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        print("gn_apply N=%d %dx%d C=%4d: %.3f ms  %.0f GB/s" % (N, H, W, C, ms, 2 * N * H * W * C * 2 / ms / 1e6))


if __name__ == "__main__":
    main()
