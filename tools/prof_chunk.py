"""Experiment: does running the high-resolution ResnetBlocks chunk-by-chunk over the batch (so that a chunk's activations
stay in the 126 MB L2 between the GroupNorm pass and the convolutions) beat one launch per layer over the whole batch?

    python tools/prof_chunk.py [C] [H] [W]

Chain per block (resnet.py:31-59, C -> C): x -gn-> a1 -conv3x3(+row)-> h -gn-> a2 -conv3x3(+x)-> y, captured in a CUDA graph;
the chunked variant reuses ONE chunk-sized a1 / h / a2 scratch for every chunk."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

nat = wsr.pkg.native
em = wsr.sub("engine")


def bview(a, i0, n):
    return em.Act(a.buf[i0:i0 + n], n, a.H, a.W, a.C, a.ld, a.coff, a.dt, a.st, a.st_off + i0 * a.st_ld, a.st_ld)


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    N = 64
    dev = torch.device("cuda:0")
    eng = em.Engine(dev, "bf16")
    torch.manual_seed(0)
    arena = em.StatsArena()
    x = eng.new_act(N, H, W, C, stats=arena)
    y = eng.new_act(N, H, W, C, stats=arena)
    hfull = eng.new_act(N, H, W, C, stats=arena)
    a1full, a2full = eng.new_act(N, H, W, C), eng.new_act(N, H, W, C)
    arena.finalize(dev)
    x.buf.copy_(torch.randn_like(x.buf, dtype=torch.float32))
    eng.gn_stats(x)
    w1 = eng.pack_conv(torch.randn(C, C, 3, 3, device=dev) / math.sqrt(9 * C), torch.randn(C, device=dev))
    w2 = eng.pack_conv(torch.randn(C, C, 3, 3, device=dev) / math.sqrt(9 * C), torch.randn(C, device=dev))
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rv = torch.randn(N, C, device=dev)
    ref = None
    for chunk in (64, 32, 16, 8, 4):
        if chunk == N:
            a1, h, a2 = a1full, hfull, a2full
        else:
            a1, a2 = eng.new_act(chunk, H, W, C), eng.new_act(chunk, H, W, C)
            h = bview(hfull, 0, chunk)          # statistics slots of the first `chunk` images, re-zeroed per chunk below

        def block():
            for i0 in range(0, N, chunk):
                xs, ys = bview(x, i0, chunk), bview(y, i0, chunk)
                if chunk != N:
                    nat.call("wsr_fill_zero", h.stats_ptr, chunk * h.st_ld * 8, eng.stream)
                eng.gn_apply(xs, g, b, 32, nat.ACT_SWISH, a1)
                eng.conv(a1, w1, h, rowvec=rv.data_ptr() + 4 * i0 * C, rowvec_ld=C)
                eng.gn_apply(h, g, b, 32, nat.ACT_SWISH, a2)
                eng.conv(a2, w2, ys, res=xs)

        nat.call("wsr_fill_zero", hfull.stats_ptr, N * hfull.st_ld * 8, eng.stream)
        nat.call("wsr_fill_zero", y.stats_ptr, N * y.st_ld * 8, eng.stream)
        block()
        torch.cuda.synchronize()
        out = y.buf.float().clone()
        if ref is None:
            ref = out
        err = float((out - ref).norm() / ref.norm())
        gr = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            block()
            torch.cuda.synchronize()
            with torch.cuda.graph(gr, stream=st):
                block()
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            gr.replay()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        fl = 2 * 2.0 * N * H * W * C * C * 9
        print("C=%d %dx%d chunk %2d: %.3f ms per block  (%.0f TFLOP/s conv-equivalent)  rel diff vs unchunked %.1e" % (C, H, W, chunk, ms, fl / ms / 1e9, err))


if __name__ == "__main__":
    main()
