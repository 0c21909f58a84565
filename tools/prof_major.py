"""K-major vs MN-major operands in the tcgen05 GEMM (is the transposed-operand path slower?).  usage: python tools/prof_major.py"""
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
    B, M, N, K = 16, 1024, 1024, 4096
    a_k = torch.randn(B, M, K, device=dev).to(torch.bfloat16)       # K-major: unit stride along k
    b_k = torch.randn(B, N, K, device=dev).to(torch.bfloat16)
    a_m = a_k.transpose(1, 2).contiguous()                          # (B, K, M): unit stride along m = MN-major
    b_m = b_k.transpose(1, 2).contiguous()
    d = torch.empty(B, M, N, device=dev, dtype=torch.bfloat16)
    ref = None
    for tag, (a, a_s), (b, b_s) in (("A K-major,  B K-major ", (a_k, (M * K, K, 1)), (b_k, (N * K, K, 1))),
                                    ("A MN-major, B K-major ", (a_m, (M * K, 1, M)), (b_k, (N * K, K, 1))),
                                    ("A K-major,  B MN-major", (a_k, (M * K, K, 1)), (b_m, (N * K, 1, N))),
                                    ("A MN-major, B MN-major", (a_m, (M * K, 1, M)), (b_m, (N * K, 1, N)))):
        def run():
            eng.gemm(a.data_ptr(), nat.BF16, a_s, b.data_ptr(), nat.BF16, b_s, d.data_ptr(), nat.BF16, (M * N, N, 1), B, M, N, K)
        run(); run()
        torch.cuda.synchronize()
        out = d.float().clone()
        if ref is None:
            ref = out
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            run()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print("%s: %.3f ms  %7.1f TFLOP/s   rel diff %.1e  (tc launches %d, simt %d)" % (tag, ms, 2.0 * B * M * N * K / ms / 1e9,
              float((out - ref).norm() / ref.norm()), eng.n_tc, eng.n_simt))


if __name__ == "__main__":
    main()
