"""Weight-gradient kernel alone: TFLOP/s per shape and batch (is the main loop or the per-launch overhead the limit?).
usage: python tools/prof_wgrad.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

nat = wsr.pkg.native
em = wsr.sub("engine")
T = wsr.sub("taps")


def main():
    dev = torch.device("cuda:0")
    eng = em.Engine(dev, "bf16")
    for (Cin, Cout, H, W) in [(64, 64, 128, 256), (128, 128, 64, 128), (256, 256, 32, 64), (512, 512, 16, 32), (512, 512, 8, 16)]:
        for N in (4, 16, 64):
            x = eng.new_act(N, H, W, Cin)
            dy = eng.new_act(N, H, W, Cout)
            x.buf.copy_(torch.randn_like(x.buf, dtype=torch.float32))
            dy.buf.copy_(torch.randn_like(dy.buf, dtype=torch.float32))
            dw = torch.zeros(Cout, Cin, 3, 3, device=dev)
            tp = T.forward_taps(3, 1, H, W)
            for _ in range(2):
                eng.wgrad(x, dy, tp, dw, (1, Cin * 9, 9), None, 1)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            reps = 10
            for _ in range(reps):
                eng.wgrad(x, dy, tp, dw, (1, Cin * 9, 9), None, 1)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            fl = 2.0 * N * H * W * Cin * Cout * 9
            print("wgrad %4d->%4d %3dx%-3d N=%2d: %.3f ms  %7.1f TFLOP/s" % (Cin, Cout, H, W, N, ms, fl / ms / 1e9))
            del x, dy


if __name__ == "__main__":
    main()
