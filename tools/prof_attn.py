"""Profiling helper: the fused attention kernel alone (for ncu / event timing).
usage: python tools/prof_attn.py B N d [reps]"""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

nat = wsr.pkg.native
Engine = wsr.sub("engine").Engine


def main():
    B, N, d = [int(v) for v in sys.argv[1:4]]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    dev = torch.device("cuda:0")
    eng = Engine(dev, "bf16")
    torch.manual_seed(0)
    H, W = N // 128, 128
    q, k, o = eng.new_act(B, H, W, d), eng.new_act(B, H, W, d), eng.new_act(B, H, W, d)
    q.buf.copy_(torch.randn_like(q.buf, dtype=torch.float32))
    k.buf.copy_(torch.randn_like(k.buf, dtype=torch.float32))
    vT = torch.randn(B, d, N, device=dev).to(torch.bfloat16)
    for _ in range(2):
        eng.attention(q, k, vT, o, None, None)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        eng.attention(q, k, vT, o, None, None)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print("attention B=%d N=%d d=%d: %.3f ms  %.1f TFLOP/s" % (B, N, d, ms, 4.0 * B * N * N * d / ms / 1e9))


if __name__ == "__main__":
    main()
