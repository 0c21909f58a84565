// attn_small_tc.cu -- fused single-head attention for SHORT key sequences and WIDE heads on tcgen05 / TMEM (sm_100a):
//   O = softmax(scale * Q K^T) V   with Nk <= 512 keys and head dimension d <= 512, one launch, scores never in HBM.
//
// Replaces, for the low-resolution levels, the reference's materialised attention (nn_modules/resnet.py:81-100 SelfAttention at
// 16x32 / 8x16: N = 512 / 128, d = 512; resdiff/guided_cross_attention.py:24-44 HF_guided_CA levels 2 and 3: N = 512, d = 256 and
// N = 128, d = 512) and this repository's own round-1 path for them (GEMM -> softmax -> GEMM: scores and probabilities went
// through HBM, 3 launches per block).  attn_tc.cu keeps the long-sequence case (N = 8192 / 2048, d = 64 / 128), where the keys are
// streamed in blocks; here the head is too wide for that design (a 128 x 512 fp32 O accumulator fills all 512 TMEM columns, leaving
// no room for score buffers), but the key sequence is short enough for the WHOLE score row block to sit in tensor memory:
//
//   phase 1  S[128 x Nk] = Q K^T            K loop over d in 64-channel chunks, {Q chunk, K chunk} ring fed by TMA, fp32 in TMEM
//   phase 2  exact single-pass softmax      row maximum over TMEM, P = exp2((S - max) * scale*log2e) -> bf16 -> 128B-swizzled smem
//   phase 3  O[128 x d] = P V               O re-uses the TMEM columns of S (dead once P is written); V^T tiles streamed by TMA
//   phase 4  O / rowsum -> bf16 -> global
//
// V comes either TRANSPOSED (vT: B x d x Nk, K-major like P) or -- v_mn -- as it leaves the q/k/v projection convolution,
// pixels x channels (B x Nk x d with a pitch): tcgen05 takes such an MN-major B operand directly (instruction-descriptor bit 16; a TMA
// box of [64 channels][64 keys] with the 128B swizzle is the canonical MN-major tile: 1 KB between 8-key atoms, 8 KB to the next 64
// channels), so the sampling plan needs no V^T GEMM at all: one convolution produces q | k | v.
// Shared memory is time-multiplexed: the phase-1 ring occupies the bytes that later hold P (<= 128 KB) and the V^T ring (3 x 32 KB).
// Warp roles (320 threads): warps 0-7 softmax + epilogue (thread = query row x column half), warp 8 TMA producer, warp 9 TMEM
// owner + single-thread MMA issuer.
#include <string.h>

#include "tc_common.cuh"

namespace wsr {

constexpr int kSmQ = 128;                       // queries per CTA
constexpr int kSmWarps = 8;
constexpr int kSmThreads = 64 + 32 * kSmWarps;
constexpr int kSmPBytes = 128 * 1024;           // P: up to 8 atoms of [128 q][64 keys] bf16
constexpr int kSmVStage = 32 * 1024;            // V^T tile: [<= 256 channels][64 keys] bf16
constexpr int kSmVStages = 3;
constexpr int kSmData = kSmPBytes + kSmVStages * kSmVStage;            // 224 KB
constexpr int kSmSmemBytes = kSmData + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*max / sum exchange*/;
constexpr int kSmMaxStages1 = 4;
static_assert(kSmSmemBytes <= 232448, "shared memory budget");

struct AttnSmallParams {
  CUtensorMap qmap, kmap, vmap;
  void* o; long long o_sb; int o_ld;      // output (B, Nq, d) bf16
  int Nk, d;
  int kpieces, kbox;                      // Nk = kpieces * kbox, kbox <= 256 (one MMA / one TMA box per piece)
  int vpieces, vbox;                      // d = vpieces * vbox, vbox <= 256
  int stages1, stage1_bytes;              // phase-1 ring
  int v_mn;                               // 1: V is (B, Nk, d) pixels x channels (MN-major B operand), 0: V^T (B, d, Nk)
  float c;                                // scale * log2(e)
};

__global__ void __launch_bounds__(kSmThreads, 1) attn_small_tc_kernel(const __grid_constant__ AttnSmallParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sP = smem;
  uint8_t* sV = smem + kSmPBytes;
  uint64_t* bars = (uint64_t*)(smem + kSmData);
  uint64_t* full1 = bars;                          // kSmMaxStages1
  uint64_t* empty1 = full1 + kSmMaxStages1;        // kSmMaxStages1
  uint64_t* v_full = empty1 + kSmMaxStages1;       // kSmVStages
  uint64_t* v_empty = v_full + kSmVStages;         // kSmVStages
  uint64_t* s_full = v_empty + kSmVStages;         // 1: all Q K^T MMAs have completed
  uint64_t* p_full = s_full + 1;                   // 1: every softmax warp has written its part of P (and is done reading S)
  uint64_t* o_full = p_full + 1;                   // 1
  uint32_t* tmem_slot = (uint32_t*)(o_full + 1);
  float* xchg = (float*)(smem + kSmData + 256);    // [2 halves][128 rows]

  constexpr int kProducerWarp = kSmWarps, kMmaWarp = kSmWarps + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kSmQ;
  const int b = blockIdx.y;
  const int nchunks = p.d >> 6;                    // 64-channel K blocks of phase 1
  const int natoms = p.Nk >> 6;                    // 64-key K blocks of phase 3
  const int S1 = p.stages1;

  if (warp == kProducerWarp && lane == 0) {
    prefetch_tmap(&p.qmap); prefetch_tmap(&p.kmap); prefetch_tmap(&p.vmap);
    for (int i = 0; i < kSmMaxStages1; ++i) { mbar_init(&full1[i], 1); mbar_init(&empty1[i], 1); }
    for (int i = 0; i < kSmVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, kSmWarps); mbar_init(o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S1; const uint32_t ph = (uint32_t)(c / S1) & 1;
        mbar_wait(&empty1[s], ph ^ 1);
        uint8_t* st = smem + s * p.stage1_bytes;
        mbar_expect_tx(&full1[s], (uint32_t)p.stage1_bytes);
        tma_load_3d(st, &p.qmap, &full1[s], c * 64, q0, b);
        for (int pc = 0; pc < p.kpieces; ++pc)
          tma_load_3d(st + 16384 + pc * p.kbox * 128, &p.kmap, &full1[s], c * 64, pc * p.kbox, b);
      }
      // the V^T ring lives in bytes the phase-1 ring was using: wait until every Q K^T MMA has read its operands
      mbar_wait(s_full, 0);
      int idx = 0;
      for (int j = 0; j < natoms; ++j)
        for (int h = 0; h < p.vpieces; ++h, ++idx) {
          const int vs = idx % kSmVStages; const uint32_t ph = (uint32_t)(idx / kSmVStages) & 1;
          mbar_wait(&v_empty[vs], ph ^ 1);
          mbar_expect_tx(&v_full[vs], (uint32_t)(p.vbox * 128));
          if (!p.v_mn) {
            tma_load_3d(sV + vs * kSmVStage, &p.vmap, &v_full[vs], j * 64, h * p.vbox, b);
          } else {
            for (int jb = 0; jb < (p.vbox >> 6); ++jb)       // [64 channels][64 keys] boxes, 8 KB apart
              tma_load_3d(sV + vs * kSmVStage + jb * 8192, &p.vmap, &v_full[vs], h * p.vbox + jb * 64, j * 64, b);
          }
        }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.kbox >> 3) << 17) | ((uint32_t)(kSmQ >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.vbox >> 3) << 17) | ((uint32_t)(kSmQ >> 4) << 24) |
                                ((uint32_t)p.v_mn << 16);
      // K-major V^T: +32 bytes per 16-key step, LBO unused; MN-major V: 16 key rows of 128 bytes per step, LBO = 8 KB (next 64 channels)
      const uint32_t v_step = p.v_mn ? 128u : 2u;
      const uint32_t v_lbo = p.v_mn ? ((8192u >> 4) << 16) : (1u << 16);
      const uint32_t kpiece_units = (uint32_t)(p.kbox * 128) >> 4;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S1; const uint32_t ph = (uint32_t)(c / S1) & 1;
        mbar_wait(&full1[s], ph);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem_u32(smem + s * p.stage1_bytes));
        const uint32_t b_lo = a_lo + (16384u >> 4);
        for (int pc = 0; pc < p.kpieces; ++pc)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_lo(tmem_base + (uint32_t)(pc * p.kbox), a_lo + (uint32_t)(kk * 2), b_lo + (uint32_t)pc * kpiece_units + (uint32_t)(kk * 2),
                         idesc_qk, (c | kk) != 0 ? 1u : 0u);
        umma_commit(&empty1[s]);
      }
      umma_commit(s_full);
      // P*V: the accumulator re-uses the score columns, so it may only start once every softmax warp has pulled its scores
      mbar_wait(p_full, 0);
      tc_fence_after();
      int idx = 0;
      for (int j = 0; j < natoms; ++j) {
        const uint32_t p_lo = desc_lo(smem_u32(sP + j * 16384));
        for (int h = 0; h < p.vpieces; ++h, ++idx) {
          const int vs = idx % kSmVStages; const uint32_t ph = (uint32_t)(idx / kSmVStages) & 1;
          mbar_wait(&v_full[vs], ph);
          tc_fence_after();
          const uint32_t v_lo = ((smem_u32(sV + vs * kSmVStage) & 0x3FFFFu) >> 4) | v_lbo;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_lo(tmem_base + (uint32_t)(h * p.vbox), p_lo + (uint32_t)(kk * 2), v_lo + (uint32_t)kk * v_step, idesc_pv, (j | kk) != 0 ? 1u : 0u);
          umma_commit(&v_empty[vs]);
        }
      }
      umma_commit(o_full);
    }
  } else {
    // ===================== softmax + epilogue (warps 0..7, thread = query row x column half) =====================
    const int quad = warp & 3;
    const int half = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const int nsc = p.Nk >> 5;                    // 32-key chunks of a score row (even: Nk % 64 == 0)
    const int sc0 = half * (nsc >> 1), sc1 = sc0 + (nsc >> 1);
    mbar_wait(s_full, 0);
    tc_fence_after();
    // exact row maximum (the whole row is in tensor memory)
    float mx = -INFINITY;
#pragma unroll 1
    for (int ch = sc0; ch < sc1; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_sel + (uint32_t)(ch * 32), v);
      float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
      for (int i = 4; i < 32; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
      }
      mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    }
    xchg[half * 128 + row] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
    mx = fmaxf(mx, xchg[(half ^ 1) * 128 + row]);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
    const float mc = mx * p.c;
    float lsum = 0.f;
#pragma unroll 1
    for (int ch = sc0; ch < sc1; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_sel + (uint32_t)(ch * 32), v);
      float e[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) e[i] = exp2f(fmaf(__uint_as_float(v[i]), p.c, -mc));
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) { s0 += e[i]; s1 += e[i + 1]; s2 += e[i + 2]; s3 += e[i + 3]; }
      lsum += (s0 + s1) + (s2 + s3);
      uint8_t* atom = sP + (ch >> 1) * 16384 + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        __nv_bfloat162* h2 = (__nv_bfloat162*)&u;
#pragma unroll
        for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(e[q * 8 + 2 * k], e[q * 8 + 2 * k + 1]);
        const int cc = (ch & 1) * 4 + q;
        *(uint4*)(atom + ((cc ^ (row & 7)) << 4)) = u;
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);
    xchg[half * 128 + row] = lsum;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
    lsum += xchg[(half ^ 1) * 128 + row];
    // epilogue: O / l -> bf16 -> global; the two halves split the d columns
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / lsum;
    __nv_bfloat16* orow = (__nv_bfloat16*)p.o + (long long)b * p.o_sb + (long long)(q0 + row) * p.o_ld;
    const int noc = p.d >> 5;                     // 32-column chunks of O (even: d % 64 == 0)
#pragma unroll 1
    for (int ch = half * (noc >> 1); ch < (half + 1) * (noc >> 1); ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_sel + (uint32_t)(ch * 32), v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        __nv_bfloat162* h2 = (__nv_bfloat162*)&u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          h2[k] = __floats2bfloat162_rn(__uint_as_float(v[q * 8 + 2 * k]) * inv, __uint_as_float(v[q * 8 + 2 * k + 1]) * inv);
        *(uint4*)(orow + ch * 32 + q * 8) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace wsr

using namespace wsr;

// 1 when wsr_attention_small_tc takes this shape
extern "C" int wsr_attention_small_tc_supported(int Nq, int Nk, int d) {
  if (Nq <= 0 || Nq % kSmQ != 0 || Nk < 64 || Nk > 512 || Nk % 64 != 0 || d < 64 || d > 512 || d % 64 != 0) return 0;
  const int kp = (Nk + 255) / 256, vp = (d + 255) / 256;
  if (Nk % kp != 0 || (Nk / kp) % 16 != 0 || d % vp != 0 || (d / vp) % 16 != 0) return 0;
  return 1;
}

static int attention_small_impl(const void* q, int q_ld, const void* k, int k_ld, const void* vT, int v_ld, void* o, int o_ld, int B,
                                int Nq, int Nk, int d, float scale, void* stream);

extern "C" int wsr_attention_small_tc(const void* q, int q_ld, const void* k, int k_ld, const void* vT, void* o, int o_ld, int B,
                                      int Nq, int Nk, int d, float scale, void* stream) {
  return attention_small_impl(q, q_ld, k, k_ld, vT, 0, o, o_ld, B, Nq, Nk, d, scale, stream);
}

extern "C" int wsr_attention_small_nhwc_tc(const void* q, int q_ld, const void* k, int k_ld, const void* v, int v_ld, void* o, int o_ld,
                                           int B, int Nq, int Nk, int d, float scale, void* stream) {
  WSR_REQUIRE(v_ld % 8 == 0 && v_ld >= d, WSR_E_UNSUPPORTED, "attention_small_nhwc_tc: v pitch");
  WSR_REQUIRE(d > 0 && (d / ((d + 255) / 256)) % 64 == 0, WSR_E_UNSUPPORTED, "attention_small_nhwc_tc: d=%d (the channel pieces must be multiples of 64)", d);
  return attention_small_impl(q, q_ld, k, k_ld, v, v_ld, o, o_ld, B, Nq, Nk, d, scale, stream);
}

// v_ld == 0: vT is V transposed (B, d, Nk); v_ld > 0: V as (B, Nk, d) with pitch v_ld
static int attention_small_impl(const void* q, int q_ld, const void* k, int k_ld, const void* vT, int v_ld, void* o, int o_ld, int B,
                                int Nq, int Nk, int d, float scale, void* stream) {
  WSR_REQUIRE(q && k && vT && o, WSR_E_INVALID, "attention_small_tc: null pointer");
  WSR_REQUIRE(B > 0 && B <= 65535 && Nq > 0 && Nk > 0, WSR_E_INVALID, "attention_small_tc: bad shape");
  WSR_REQUIRE(wsr_attention_small_tc_supported(Nq, Nk, d), WSR_E_UNSUPPORTED,
              "attention_small_tc: needs Nq %% 128 == 0, 64 <= Nk <= 512 with Nk %% 64 == 0, 64 <= d <= 512 with d %% 64 == 0 (got Nq=%d Nk=%d d=%d)", Nq, Nk, d);
  WSR_REQUIRE(q_ld % 8 == 0 && k_ld % 8 == 0 && o_ld % 8 == 0 && q_ld >= d && k_ld >= d && o_ld >= d, WSR_E_UNSUPPORTED, "attention_small_tc: pitches");
  WSR_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)vT | (uintptr_t)o) & 15) == 0, WSR_E_UNSUPPORTED, "attention_small_tc: 16-byte alignment");
  AttnSmallParams p;
  memset(&p, 0, sizeof(p));
  p.Nk = Nk; p.d = d;
  p.kpieces = (Nk + 255) / 256; p.kbox = Nk / p.kpieces;
  p.vpieces = (d + 255) / 256; p.vbox = d / p.vpieces;
  p.stage1_bytes = 16384 + Nk * 128;
  p.stages1 = kSmData / p.stage1_bytes;
  if (p.stages1 > kSmMaxStages1) p.stages1 = kSmMaxStages1;
  WSR_REQUIRE(p.stages1 >= 1, WSR_E_UNSUPPORTED, "attention_small_tc: stage size");
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)Nq, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)q_ld * 2, (uint64_t)Nq * q_ld * 2};
    uint32_t box[3] = {64, (uint32_t)kSmQ, 1};
    if ((rc = encode_map(&p.qmap, q, 3, dims, str, box))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)Nk, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)k_ld * 2, (uint64_t)Nk * k_ld * 2};
    uint32_t box[3] = {64, (uint32_t)p.kbox, 1};
    if ((rc = encode_map(&p.kmap, k, 3, dims, str, box))) return rc;
  }
  p.v_mn = v_ld > 0 ? 1 : 0;
  if (!p.v_mn) {
    uint64_t dims[3] = {(uint64_t)Nk, (uint64_t)d, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)Nk * 2, (uint64_t)d * Nk * 2};
    uint32_t box[3] = {64, (uint32_t)p.vbox, 1};
    if ((rc = encode_map(&p.vmap, vT, 3, dims, str, box))) return rc;
  } else {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)Nk, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)v_ld * 2, (uint64_t)Nk * v_ld * 2};
    uint32_t box[3] = {64, 64, 1};
    if ((rc = encode_map(&p.vmap, vT, 3, dims, str, box))) return rc;
  }
  p.o = o; p.o_sb = (long long)Nq * o_ld; p.o_ld = o_ld;
  p.c = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    WSR_CUDA_OK(cudaFuncSetAttribute(attn_small_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmSmemBytes));
    attr_set = true;
  }
  WSR_CUDA_OK(launch_pdl(attn_small_tc_kernel, dim3(Nq / kSmQ, B), dim3(kSmThreads), (size_t)kSmSmemBytes, (cudaStream_t)stream, p));
  return WSR_OK;
}
