// tma_probe.cu -- how fast can ONE SM pull 128-row x 128-byte operand tiles out of L2 with TMA, and does cluster multicast change it?
//
// The deep UNet levels at small batch are bound by operand fetch, not by MMA (with the MMAs switched off a
// 512->512 convolution on 8x16 images takes 21 of its 24 us).  This probe isolates the fetch: every CTA streams `iters` tiles of
// [128 rows][64 bf16] (16 KB, 128B swizzle, the A operand of one K block) through an 8-stage ring, a second thread consumes
// them (wait full -> release), nothing else runs.  Variants:
//   pitch    : distance between consecutive rows of a tile in bytes -- 1024 (NHWC with 512 channels: every row of the box is its own
//              128-byte segment) or 128 (the 16 KB of a tile are contiguous)
//   cluster  : 1 = every CTA loads whole tiles; 2 / 4 / 8 = the CTAs of a cluster want the SAME tile (column tiles sharing an
//              activation tile), each loads 128/cs rows and multicasts them to all
//   grid     : number of CTAs (one per SM)
// build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu -lcuda
// prints GB/s delivered per SM and in aggregate.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kStages = 8;
constexpr int kTile = 16384;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint64_t* b, uint32_t rank) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(s32(b)), "r"(rank) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)), "l"(m),
               "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma3_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(s32(dst)),
      "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct Params {
  CUtensorMap map;      // {channels, 128 rows, tiles}, box {64, 128 / cs, 1}
  int iters, cs, ntiles, nchunks, multicast;
};

__global__ void __launch_bounds__(64, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + kStages * kTile);
  uint64_t* empty = full + kStages;
  const uint32_t rank = p.cs > 1 ? ctarank() : 0;
  const int cluster_id = blockIdx.x / p.cs;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], p.multicast ? p.cs : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (p.cs > 1) cluster_sync(); else __syncthreads();
  const int rows = 128 / (p.multicast ? p.cs : 1);
  if (threadIdx.x == 0) {
    // producer
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < p.iters; ++it) {
      const int tile = (cluster_id * 7 + it / p.nchunks) % p.ntiles, c0 = (it % p.nchunks) * 64;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect(&full[s], kTile);
      if (p.multicast) tma3_mc(smem + s * kTile + rank * rows * 128, &p.map, &full[s], c0, rank * rows, tile, (uint16_t)((1u << p.cs) - 1));
      else tma3(smem + s * kTile, &p.map, &full[s], c0, 0, tile);
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    // consumer
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(&full[s], ph);
      if (p.multicast) { for (int r = 0; r < p.cs; ++r) arrive_remote(&empty[s], r); }
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
  }
  __syncwarp();
  if (p.cs > 1) cluster_sync(); else __syncthreads();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2048;
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  const int ntiles = 64;
  const size_t bytes = (size_t)ntiles * 128 * 1024;   // 8 MB: L2-resident either way
  void* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kTile + 2048));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("%-8s %-4s %-4s %-5s %10s %12s %12s\n", "pitch", "cs", "mc", "grid", "us", "GB/s per SM", "GB/s total");
  for (int pitch : {1024, 128}) {
    const int nchunks = pitch / 128;
    for (int cfg = 0; cfg < 7; ++cfg) {
      const int cs = cfg == 0 ? 1 : (cfg <= 3 ? (1 << cfg) : (1 << (cfg - 3)));
      const int mc = cfg >= 4 ? 1 : 0;                 // cfg 1..3: clusters WITHOUT multicast (each CTA loads the whole shared tile itself)
      for (int grid : {8, 16, 64, 144}) {
        if (grid > nsm) continue;
        Params p; memset(&p, 0, sizeof(p));
        cuuint64_t dims[3] = {(cuuint64_t)(pitch / 2), 128, (cuuint64_t)ntiles};
        cuuint64_t str[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * 128};
        cuuint32_t box[3] = {64, (cuuint32_t)(128 / (mc ? cs : 1)), 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        p.iters = iters; p.cs = cs; p.ntiles = ntiles; p.nchunks = nchunks; p.multicast = mc;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid); lc.blockDim = dim3(64); lc.dynamicSmemBytes = kStages * kTile + 2048;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = cs > 1 ? 1 : 0;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          CK(cudaEventRecord(e0));
          cudaError_t le = cudaLaunchKernelEx(&lc, probe_kernel, p);
          if (le != cudaSuccess) { printf("launch failed (pitch %d cs %d grid %d): %s\n", pitch, cs, grid, cudaGetErrorString(le)); cudaGetLastError(); best = -1.f; break; }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0 && ms < best) best = ms;
        }
        if (best < 0) continue;
        const double per_sm = (double)iters * kTile / (best * 1e-3) / 1e9;
        printf("%-8d %-4d %-4d %-5d %10.1f %12.1f %12.1f\n", pitch, cs, mc, grid, best * 1e3, per_sm, per_sm * grid);
        fflush(stdout);
      }
    }
  }
  return 0;
}
