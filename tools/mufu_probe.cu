// mufu_probe.cu -- special-function throughput per SM on sm_100a: ex2.approx.ftz.f32 vs the packed ex2.approx.ftz.f16x2 / .bf16x2
// (two exponentials per lane and instruction).  The N = 8192 attention kernel is bound by the exponentials of its softmax (ncu: XU pipe
// 60 % busy, tensor pipe 30 %); if the packed forms issue at the same rate they double its ceiling.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/mufu_probe tools/mufu_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, int iters, uint32_t seed) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;      // bit patterns of small negative numbers in every format
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 3) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; ex2.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(r[i]));
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  uint32_t* out; cudaMalloc(&out, (size_t)nsm * 8 * 256 * 4);
  const int iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "ex2.approx.f16 (scalar)"};
  for (int mode = 0; mode < 4; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<nsm * 8, 256>>>(out, iters, 0xBC00BC00u);
      if (mode == 1) probe<1><<<nsm * 8, 256>>>(out, iters, 0xBC00BC00u);
      if (mode == 2) probe<2><<<nsm * 8, 256>>>(out, iters, 0xBF80BF80u);
      if (mode == 3) probe<3><<<nsm * 8, 256>>>(out, iters, 0xBC00BC00u);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double instr = (double)nsm * 8 * 256 * iters * 8;                // lane-instructions
    const double per_clk_sm = instr / (best * 1e-3) / ((double)khz * 1e3) / nsm;
    printf("%-30s %8.3f ms  %6.2f lane-instructions / clk / SM (at the nominal %d MHz)  -> %6.2f exponentials / clk / SM\n", names[mode], best, per_clk_sm,
           khz / 1000, per_clk_sm * ((mode == 1 || mode == 2) ? 2 : 1));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
