// attn_tc.cu -- fused single-head attention  O = softmax(scale * Q K^T) V  on tcgen05 / TMEM (sm_100a).
//
// Replaces the reference's materialised attention (nn_modules/resnet.py:90-97, guided_cross_attention.py:34-41:
// einsum -> (B,N,N) fp32 matrix -> .contiguous() -> /sqrt(C) -> softmax -> einsum), whose N = 8192 instance writes
// 268 MB per sample per step.  Here the score matrix never leaves the SM.
//
// One CTA owns 128 query rows and streams the keys in blocks of 128, in two passes:
//   pass A:  S = Q K^T (tcgen05.mma into TMEM) for a few SAMPLED key blocks (4, evenly spread), row maxima only
//   pass B:  S for every block, P = exp2((S - ref) * scale*log2e) -> bf16 -> swizzled smem, O += P V, row sums
// Softmax is invariant to the shift `ref`, which only has to keep the exponentials inside the floating-point range;
// it need not be the exact row maximum.  `ref` = the maximum over the sampled keys is an actual score of the row, so
// nothing that matters can underflow, and a score above it simply gives P > 1 (fp32 sums, bf16 P and the fp32 O
// accumulator all have 8 exponent bits; the exponent is clamped at +64 as a safety net, which cannot trigger unless a
// row's scores differ by more than 44 natural units between the sampled and the unsampled keys).  Knowing the shift
// before pass B removes the online-softmax rescaling of the O accumulator (and the dependency it creates between
// softmax and the previous P*V); sampling removes 15/16 of the pass-A tensor and TMEM-read work of the N = 8192 layer.
// Up to 4 key blocks the maximum is exact.
// V is consumed TRANSPOSED (vT: B x d x Nk, produced that way by the K/V projection GEMM) so that both operands of P*V
// are K-major and can be fed by plain 128B-swizzled TMA boxes.
//
// Warp roles (320 threads): warps 0-7 softmax + epilogue, warp 8 TMA producer, warp 9 TMEM owner + MMA issuer
// (thread = query row x key half; tcgen05.ld 32x32b gives each thread 32 consecutive keys of its row; two warps per
// SMSP keep the SFU busy while the other converts / stores).
#include <string.h>

#include "tc_common.cuh"

namespace wsr {

constexpr int kBQ = 128;     // queries per CTA
constexpr int kBK = 128;     // keys per block
constexpr int kSoftmaxWarps = 8;                        // two warps per TMEM lane quadrant, each owning 64 of the 128 keys
constexpr int kAttnThreads = 64 + 32 * kSoftmaxWarps;

template <int D> struct AttnCfg {
  static constexpr int kChunks = D / 64;                  // 64-channel K chunks of Q / K
  static constexpr int kQBytes = kChunks * 16384;
  static constexpr int kKBytes = kChunks * 16384;         // one key block [128 keys][D]
  static constexpr int kVAtom = D * 128;                  // [D rows][64 keys] bf16
  static constexpr int kVBytes = 2 * kVAtom;              // one key block of V^T
  static constexpr int kPBytes = 2 * 16384;               // [128 q][128 keys] bf16 as two 64-key atoms
  static constexpr int kKStages = D == 64 ? 3 : 2;
  static constexpr int kVStages = D == 64 ? 3 : 2;
  static constexpr int kSmemData = kQBytes + kKStages * kKBytes + kVStages * kVBytes + 2 * kPBytes;
  static constexpr int kSmemBytes = kSmemData + 1024 + 256 + 1024;   // + align slack + barriers + max/sum exchange
  static constexpr uint32_t kIdescQK = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBK >> 3) << 17) | ((uint32_t)(kBQ >> 4) << 24);
  static constexpr uint32_t kIdescPV = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(kBQ >> 4) << 24);
};

struct AttnParams {
  CUtensorMap qmap, kmap, vmap;
  void* o; long long o_sb; int o_ld;      // output (B, Nq, d) bf16
  int nblocks;                            // Nk / 128
  int na, stride_a;                       // pass A: na sampled key blocks, block g*stride_a
  float c;                                // scale * log2(e)
};

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int KS = Cfg::kKStages, VS = Cfg::kVStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kQBytes;
  uint8_t* sV = sK + KS * Cfg::kKBytes;
  uint8_t* sP = sV + VS * Cfg::kVBytes;
  uint64_t* bars = (uint64_t*)(smem + Cfg::kSmemData);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = q_full + 1;           // KS
  uint64_t* k_empty = k_full + KS;         // KS
  uint64_t* v_full = k_empty + KS;         // VS
  uint64_t* v_empty = v_full + VS;         // VS
  uint64_t* s_full = v_empty + VS;         // 2
  uint64_t* s_empty = s_full + 2;          // 2
  uint64_t* p_full = s_empty + 2;          // 2
  uint64_t* p_empty = p_full + 2;          // 2
  uint64_t* o_full = p_empty + 2;          // 1
  uint32_t* tmem_slot = (uint32_t*)(o_full + 1);
  float* xchg = (float*)(tmem_slot + 2);    // [2 halves][128 rows]: row max / row sum exchange between the two key halves

  // warp roles: 0..7 softmax/epilogue, 8 TMA producer, 9 MMA issuer (highest warp id = highest issue priority)
  constexpr int kProducerWarp = kSoftmaxWarps, kMmaWarp = kSoftmaxWarps + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBQ;
  const int b = blockIdx.y;
  const int NB = p.nblocks, NA = p.na;

  if (warp == kProducerWarp && lane == 0) {
    prefetch_tmap(&p.qmap); prefetch_tmap(&p.kmap); prefetch_tmap(&p.vmap);
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], kSoftmaxWarps); mbar_init(&p_full[i], kSoftmaxWarps); mbar_init(&p_empty[i], 1); }
    mbar_init(o_full, 1);
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
  const uint32_t tmem_S = tmem_base;            // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;      // D columns

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int c = 0; c < Cfg::kChunks; ++c) tma_load_3d(sQ + c * 16384, &p.qmap, q_full, c * 64, q0, b);
      int vj = 0;
      for (int g = 0; g < NA + NB; ++g) {
        const int j = g < NA ? g * p.stride_a : g - NA;
        const int ks = g % KS; const uint32_t kph = (uint32_t)(g / KS) & 1;
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_expect_tx(&k_full[ks], Cfg::kKBytes);
        for (int c = 0; c < Cfg::kChunks; ++c) tma_load_3d(sK + ks * Cfg::kKBytes + c * 16384, &p.kmap, &k_full[ks], c * 64, j * kBK, b);
        if (g >= NA) {
          const int vs = vj % VS; const uint32_t vph = (uint32_t)(vj / VS) & 1;
          mbar_wait(&v_empty[vs], vph ^ 1);
          mbar_expect_tx(&v_full[vs], Cfg::kVBytes);
          tma_load_3d(sV + vs * Cfg::kVBytes, &p.vmap, &v_full[vs], j * kBK, 0, b);
          tma_load_3d(sV + vs * Cfg::kVBytes + Cfg::kVAtom, &p.vmap, &v_full[vs], j * kBK + 64, 0, b);
          ++vj;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_wait(q_full, 0);
      tc_fence_after();
      const uint32_t q_lo = desc_lo(smem_u32(sQ));
      auto issue_qk = [&](int g) {
        const int ks = g % KS; const uint32_t kph = (uint32_t)(g / KS) & 1;
        const int sb = g & 1; const uint32_t sph = (uint32_t)(g >> 1) & 1;
        mbar_wait(&s_empty[sb], sph ^ 1);
        mbar_wait(&k_full[ks], kph);
        tc_fence_after();
        // lean issue path (32-bit descriptor low words, see tc_common.cuh): on this single-thread dependent chain every
        // integer instruction costs ~5 cycles, and rebuilding two 64-bit descriptors per MMA made the ISSUE the bottleneck
        const uint32_t k_lo = desc_lo(smem_u32(sK + ks * Cfg::kKBytes));
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_lo(tmem_S + (uint32_t)(sb * kBK), q_lo + (uint32_t)(c * 1024 + kk * 2), k_lo + (uint32_t)(c * 1024 + kk * 2),
                         Cfg::kIdescQK, (c | kk) != 0 ? 1u : 0u);
        umma_commit(&k_empty[ks]);
        umma_commit(&s_full[sb]);
      };
      // pass A: scores of the sampled blocks only
      for (int g = 0; g < NA; ++g) issue_qk(g);
      // pass B: scores TWO blocks ahead of P*V.  P*V(j) can only be issued once softmax(j) has delivered P(j); with the
      // scores just one block ahead, Q K^T(j+2) sat behind that wait and softmax(j+2) could not start before softmax(j+1)
      // had finished AND Q K^T(j+2) had run (period = softmax + Q K^T).  The score buffer of block j is free as soon as
      // softmax(j) has pulled it into registers, so Q K^T(j+2) is issued right after P*V(j): softmax(j+1) overlaps both.
      issue_qk(NA);
      if (NB > 1) issue_qk(NA + 1);
      for (int j = 0; j < NB; ++j) {
        const int pb = j & 1; const uint32_t pph = (uint32_t)(j >> 1) & 1;
        const int vs = j % VS; const uint32_t vph = (uint32_t)(j / VS) & 1;
        mbar_wait(&p_full[pb], pph);
        mbar_wait(&v_full[vs], vph);
        tc_fence_after();
        const uint32_t p_lo = desc_lo(smem_u32(sP + pb * Cfg::kPBytes));
        const uint32_t v_lo = desc_lo(smem_u32(sV + vs * Cfg::kVBytes));
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_lo(tmem_O, p_lo + (uint32_t)(a * 1024 + kk * 2), v_lo + (uint32_t)(a * (Cfg::kVAtom >> 4) + kk * 2),
                         Cfg::kIdescPV, (j | a | kk) != 0 ? 1u : 0u);
        umma_commit(&p_empty[pb]);
        umma_commit(&v_empty[vs]);
        if (j + 2 < NB) issue_qk(NA + j + 2);
      }
      umma_commit(o_full);
    }
  } else {
    // ===================== softmax + epilogue (warps 0..7, thread = query row x key half) =====================
    const int quad = warp & 3;
    const int half = warp >> 2;          // 0: keys [0,64) of every block, 1: keys [64,128)
    const int row = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    float mx = -INFINITY;
    // pass A: row maximum of the raw scores of the sampled key blocks
    for (int g = 0; g < NA; ++g) {
      const int sb = g & 1; const uint32_t sph = (uint32_t)(g >> 1) & 1;
      mbar_wait(&s_full[sb], sph);
      tc_fence_after();
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_sel + (uint32_t)(sb * kBK + half * 64 + c2 * 32), v);
        // four independent max chains (a single running max would be a 64-deep dependent chain per block)
        float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
        mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
    }
    xchg[half * 128 + row] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSoftmaxWarps) : "memory");
    mx = fmaxf(mx, xchg[(half ^ 1) * 128 + row]);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSoftmaxWarps) : "memory");
    const float mc = mx * p.c;
    float lsum = 0.f;
    // pass B: probabilities -> smem (K-major, 128B swizzle), row sums
    for (int j = 0; j < NB; ++j) {
      const int g = NA + j;
      const int sb = g & 1; const uint32_t sph = (uint32_t)(g >> 1) & 1;
      const int pb = j & 1; const uint32_t pph = (uint32_t)(j >> 1) & 1;
      mbar_wait(&s_full[sb], sph);
      tc_fence_after();
      uint8_t* atom = sP + pb * Cfg::kPBytes + half * 16384 + row * 128;     // this half = one 64-key atom of P
      // both 32-key chunks are fetched up front; once they sit in registers the score buffer goes back to the MMA issuer
      // (the next Q K^T starts while the exponentials are still being computed), and the fully unrolled body lets the
      // compiler interleave the MUFU stream of one chunk with the sums / conversions / stores of the other
      uint32_t v[2][32];
      tmem_ld32_nowait(tmem_S + lane_sel + (uint32_t)(sb * kBK + half * 64), v[0]);
      tmem_ld32_nowait(tmem_S + lane_sel + (uint32_t)(sb * kBK + half * 64 + 32), v[1]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      mbar_wait(&p_empty[pb], pph ^ 1);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = exp2f(fminf(fmaf(__uint_as_float(v[c2][i]), p.c, -mc), 64.f));
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;       // independent partial sums
#pragma unroll
        for (int i = 0; i < 32; i += 4) { s0 += e[i]; s1 += e[i + 1]; s2 += e[i + 2]; s3 += e[i + 3]; }
        lsum += (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          __nv_bfloat162* h = (__nv_bfloat162*)&u;
#pragma unroll
          for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(e[q * 8 + 2 * k], e[q * 8 + 2 * k + 1]);
          const int cc = c2 * 4 + q;
          *(uint4*)(atom + ((cc ^ (row & 7)) << 4)) = u;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[pb]);
    }
    xchg[half * 128 + row] = lsum;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSoftmaxWarps) : "memory");
    lsum += xchg[(half ^ 1) * 128 + row];
    // epilogue: O / l -> bf16 -> global; the two halves split the D columns
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / lsum;
    __nv_bfloat16* orow = (__nv_bfloat16*)p.o + (long long)b * p.o_sb + (long long)(q0 + row) * p.o_ld;
#pragma unroll 1
    for (int c2 = 0; c2 < D / 64; ++c2) {
      const int ch = half * (D / 64) + c2;
      uint32_t v[32];
      tmem_ld32(tmem_O + lane_sel + (uint32_t)(ch * 32), v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        __nv_bfloat162* h = (__nv_bfloat162*)&u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          h[k] = __floats2bfloat162_rn(__uint_as_float(v[q * 8 + 2 * k]) * inv, __uint_as_float(v[q * 8 + 2 * k + 1]) * inv);
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

template <int D>
static int launch_attn(const AttnParams& p, int B, int Nq, cudaStream_t st) {
  using Cfg = AttnCfg<D>;
  static bool attr_set = false;
  if (!attr_set) {
    WSR_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid(Nq / kBQ, B);
  attn_tc_kernel<D><<<grid, kAttnThreads, Cfg::kSmemBytes, st>>>(p);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

}  // namespace wsr

using namespace wsr;

extern "C" int wsr_attention_tc(const void* q, int q_ld, const void* k, int k_ld, const void* vT, void* o, int o_ld, int B,
                                int Nq, int Nk, int d, float scale, void* stream) {
  WSR_REQUIRE(q && k && vT && o, WSR_E_INVALID, "attention_tc: null pointer");
  WSR_REQUIRE(B > 0 && B <= 65535 && Nq > 0 && Nk > 0, WSR_E_INVALID, "attention_tc: bad shape");
  WSR_REQUIRE(d == 64 || d == 128, WSR_E_UNSUPPORTED, "attention_tc: head dim %d (only 64, 128)", d);
  WSR_REQUIRE(Nq % kBQ == 0 && Nk % kBK == 0, WSR_E_UNSUPPORTED, "attention_tc: Nq, Nk must be multiples of 128 (got %d, %d)", Nq, Nk);
  WSR_REQUIRE(q_ld % 8 == 0 && k_ld % 8 == 0 && o_ld % 8 == 0 && q_ld >= d && k_ld >= d && o_ld >= d, WSR_E_UNSUPPORTED, "attention_tc: pitches");
  WSR_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)vT | (uintptr_t)o) & 15) == 0, WSR_E_UNSUPPORTED, "attention_tc: 16-byte alignment");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)Nq, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)q_ld * 2, (uint64_t)Nq * q_ld * 2};
    uint32_t box[3] = {64, (uint32_t)kBQ, 1};
    if ((rc = encode_map(&p.qmap, q, 3, dims, str, box))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)Nk, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)k_ld * 2, (uint64_t)Nk * k_ld * 2};
    uint32_t box[3] = {64, (uint32_t)kBK, 1};
    if ((rc = encode_map(&p.kmap, k, 3, dims, str, box))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Nk, (uint64_t)d, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)Nk * 2, (uint64_t)d * Nk * 2};
    uint32_t box[3] = {64, (uint32_t)d, 1};
    if ((rc = encode_map(&p.vmap, vT, 3, dims, str, box))) return rc;
  }
  p.o = o; p.o_sb = (long long)Nq * o_ld; p.o_ld = o_ld;
  p.nblocks = Nk / kBK;
  p.na = p.nblocks < 4 ? p.nblocks : 4;
  p.stride_a = p.nblocks / p.na;
  p.c = scale * 1.4426950408889634f;
  cudaStream_t st = (cudaStream_t)stream;
  return d == 64 ? launch_attn<64>(p, B, Nq, st) : launch_attn<128>(p, B, Nq, st);
}
