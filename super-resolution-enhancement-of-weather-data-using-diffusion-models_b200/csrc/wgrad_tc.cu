// wgrad_tc.cu -- weight gradient of the tap-table convolutions on tcgen05 / TMEM (sm_100a).
//
//   dw[tap][co][ci] += sum_{pixels p} dy[p][co] * X[p + shift(tap)][ci]
//
// Per tap this is a GEMM  D[co][ci] = A[co][p] * B[ci][p]^T  whose reduction dimension is the PIXEL index, while both
// operands are NHWC activations, i.e. contiguous along channels = along M / N.  tcgen05.mma takes such "MN-major"
// operands directly: the instruction descriptor's a_major / b_major bits select them, and a TMA box of
// [64 channels][64 pixels] with the 128-byte swizzle IS the canonical MN-major SWIZZLE_128B tile (one 128-byte row per
// pixel, 8-pixel atoms 1024 bytes apart = stride byte offset; the next 64 channels are a second box 8 KB further =
// leading byte offset).  No transposed copy of any activation is ever made.  The convolution's zero padding is again
// the TMA out-of-bounds fill: the X box of tap (dy, dx) is the dY box shifted by (dy, dx).
//
// Work decomposition: one CTA = (tap, 128-channel Cout tile, BLOCK_N-channel Cin tile, K split).  The K range (pixel
// tiles of 64 pixels x segments) is split so that about one CTA per SM exists; partial tiles are added to dw with fp32
// reductions (red.global.add), dw being zeroed once per step by the caller.  Strided output addressing lets the result
// land directly in the reference's OIHW parameter layout.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM owner + single-thread MMA issuer, warps 2-5 epilogue.
// Replaces the weight half of `l_pix.backward()` (models/diffusion_models/model.py:67) for every nn.Conv2d on the path.
#include <cuda.h>
#include <string.h>

#include "tc_common.cuh"

namespace wsr {

constexpr int kWgSub = 64 * 128;        // one [64 pixels][64 channels] bf16 sub-block = 8 KB
constexpr int kWgMaxSeg = 4;
constexpr int kWgThreads = 192;

struct WgSeg { int16_t amap, bmap, dx, dy; };

struct WgParams {
  CUtensorMap amap[4];                  // dY views: plain, or the 4 phase-subsampled views (upsample convolution)
  CUtensorMap bmap[4];                  // X views: plain, or the 4 phase-subsampled views (stride-2 convolution)
  WgSeg seg[WSR_MAX_TAPS][kWgMaxSeg];
  int wtap[WSR_MAX_TAPS];
  int ntaps, nseg;
  int t1, t2, t3, g1, g2, g3;           // pixel tile (t1*t2*t3 == 64) and number of tiles along (w, h, n)
  int m_tiles, n_tiles, splits;
  int Cout, Cin;
  float* dw; long long stap, sco, sci;
};

template <int BLOCK_N> struct WgCfg {
  static constexpr int kABytes = 2 * kWgSub;
  static constexpr int kBBytes = (BLOCK_N / 64) * kWgSub;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N >= 256 ? 4 : (BLOCK_N >= 128 ? 6 : 8);
  static constexpr int kSmemData = kStages * kStageBytes;
  static constexpr int kSmemBytes = kSmemData + 1024 + 256;
  // D = f32, A = B = bf16, both MN-major (bits 15, 16), N, M = 128
  static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

// MN-major SWIZZLE_128B operand descriptor: LBO = 8 KB (next 64 channels), SBO = 1 KB (next 8 pixels)
constexpr uint32_t kWgDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t wg_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | ((uint32_t)(kWgSub >> 4) << 16); }
__device__ __forceinline__ void wg_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(kWgDescHi), "r"(idesc), "r"(accumulate) : "memory");
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  using Cfg = WgCfg<BLOCK_N>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + Cfg::kSmemData);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint32_t* tmem_slot = (uint32_t*)(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // decode the work item
  int w = blockIdx.x;
  const int split = w % p.splits; w /= p.splits;
  const int nt = w % p.n_tiles; w /= p.n_tiles;
  const int mt = w % p.m_tiles;
  const int tap = w / p.m_tiles;
  const int nptiles = p.g1 * p.g2 * p.g3;
  const int total_kb = p.nseg * nptiles;
  const int kb0 = (int)((long long)split * total_kb / p.splits);
  const int kb1 = (int)((long long)(split + 1) * total_kb / p.splits);
  const int co0 = mt * 128, ci0 = nt * BLOCK_N;
  const int a_blocks = (p.Cout - co0) > 64 ? 2 : 1;
  int b_blocks = (p.Cin - ci0 + 63) / 64;
  if (b_blocks > BLOCK_N / 64) b_blocks = BLOCK_N / 64;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) { prefetch_tmap(&p.amap[i]); prefetch_tmap(&p.bmap[i]); }
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BLOCK_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // blocks of the B tile that lie beyond Cin are never loaded: clear them once so that the MMA reads zeros, not stale
  // shared memory (NaN bit patterns would poison whole accumulator rows through 0 * NaN)
  if (b_blocks < BLOCK_N / 64 || a_blocks < 2) {
    for (int s = 0; s < S; ++s) {
      uint4* st = (uint4*)(smem + s * Cfg::kStageBytes);
      const uint4 z = make_uint4(0, 0, 0, 0);
      if (a_blocks < 2)
        for (int i = threadIdx.x; i < kWgSub / 16; i += kWgThreads) st[kWgSub / 16 + i] = z;
      for (int i = threadIdx.x + b_blocks * (kWgSub / 16); i < (BLOCK_N / 64) * (kWgSub / 16); i += kWgThreads) st[Cfg::kABytes / 16 + i] = z;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int sg = kb / nptiles;
        int pt = kb - sg * nptiles;
        const int i1 = pt % p.g1; pt /= p.g1;
        const int i2 = pt % p.g2;
        const int i3 = pt / p.g2;
        const WgSeg e = p.seg[tap][sg];
        const int c1 = i1 * p.t1, c2 = i2 * p.t2, c3 = i3 * p.t3;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * Cfg::kStageBytes;
        mbar_expect_tx(&full[s], (uint32_t)((a_blocks + b_blocks) * kWgSub));
        for (int j = 0; j < a_blocks; ++j) tma_load_4d(st + j * kWgSub, &p.amap[e.amap], &full[s], co0 + 64 * j, c1, c2, c3);
        for (int j = 0; j < b_blocks; ++j)
          tma_load_4d(st + Cfg::kABytes + j * kWgSub, &p.bmap[e.bmap], &full[s], ci0 + 64 * j, c1 + e.dx, c2 + e.dy, c3);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = wg_desc_lo(smem_u32(smem + s * Cfg::kStageBytes));
        const uint32_t b_lo = wg_desc_lo(smem_u32(smem + s * Cfg::kStageBytes + Cfg::kABytes));
        // 16 pixels per MMA = 16 rows of 128 bytes = 2048 bytes = 128 address units
#pragma unroll
        for (int k = 0; k < 4; ++k) wg_umma(tmem_base, a_lo + 128u * k, b_lo + 128u * k, Cfg::kIdesc, (kb != kb0 || k != 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull);
    }
  } else if (kb1 > kb0) {
    // ===================== epilogue (warps 2..5): TMEM -> fp32 reductions into dw =====================
    const int quad = warp & 3;
    const int co = co0 + quad * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    float* base = p.dw + (long long)p.wtap[tap] * p.stap + (long long)co * p.sco;
#pragma unroll 1
    for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ch * 32), v);
      if (co < p.Cout) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int ci = ci0 + ch * 32 + j;
          if (ci < p.Cin) atomicAdd(base + (long long)ci * p.sci, __uint_as_float(v[j]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N) : "memory");
  }
}

template <int BLOCK_N>
static int launch_wgrad(const WgParams& p, cudaStream_t st) {
  using Cfg = WgCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    WSR_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int grid = p.ntaps * p.m_tiles * p.n_tiles * p.splits;
  wgrad_tc_kernel<BLOCK_N><<<grid, kWgThreads, Cfg::kSmemBytes, st>>>(p);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

int validate_taps(const WsrTapTable* t);

// per-channel sums of an NHWC tensor: out[c] += sum_p x[p][c]   (bias gradients)
template <typename T>
__global__ void __launch_bounds__(256) col_sums_kernel(const T* __restrict__ x, int64_t pixels, int C, int ld, int64_t chunk, float* out) {
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  const int64_t p1 = p0 + chunk < pixels ? p0 + chunk : pixels;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    int64_t p = p0;
    for (; p + 1 < p1; p += 2) { s0 += ldf<T>(x + p * ld + c); s1 += ldf<T>(x + (p + 1) * ld + c); }
    if (p < p1) s0 += ldf<T>(x + p * ld + c);
    atomicAdd(out + c, s0 + s1);
  }
}

}  // namespace wsr

using namespace wsr;

static inline int wg_cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

extern "C" int wsr_col_sums(const void* x, int x_dtype, int64_t pixels, int C, int x_ld, float* out, void* stream) {
  WSR_REQUIRE(x && out && valid_dtype(x_dtype) && pixels > 0 && C > 0 && x_ld >= C, WSR_E_INVALID, "col_sums: bad argument");
  int64_t chunk = (pixels + 591) / 592;
  if (chunk < 16) chunk = 16;
  const unsigned blocks = (unsigned)((pixels + chunk - 1) / chunk);
  const int threads = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  if (x_dtype == WSR_BF16) col_sums_kernel<__nv_bfloat16><<<blocks, threads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, pixels, C, x_ld, chunk, out);
  else col_sums_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>((const float*)x, pixels, C, x_ld, chunk, out);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_conv_wgrad_tc(const WsrWgradDesc* d, const WsrTapTable* t, void* stream) {
  WSR_REQUIRE(d && t, WSR_E_INVALID, "wgrad_tc: null descriptor");
  WSR_REQUIRE(d->x && d->dy && d->dw, WSR_E_INVALID, "wgrad_tc: null x/dy/dw");
  int rc = validate_taps(t);
  if (rc) return rc;
  WSR_REQUIRE(t->in_sub <= 2, WSR_E_UNSUPPORTED, "wgrad_tc: in_sub=%d (1 or 2)", t->in_sub);
  WSR_REQUIRE(d->x_dtype == WSR_BF16 && d->dy_dtype == WSR_BF16, WSR_E_UNSUPPORTED, "wgrad_tc: bf16 operands only");
  WSR_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && d->x_ld >= d->Cin && d->dy_ld >= d->Cout, WSR_E_INVALID, "wgrad_tc: bad shape");
  WSR_REQUIRE(d->x_ld % 8 == 0 && d->dy_ld % 8 == 0 && (((uintptr_t)d->x) & 15) == 0 && (((uintptr_t)d->dy) & 15) == 0, WSR_E_UNSUPPORTED,
              "wgrad_tc: pitches %% 8 and 16-byte aligned bases required");
  WSR_REQUIRE(t->out_mul == 1 && t->out_py == 0 && t->out_px == 0 && t->GH == t->OH && t->GW == t->OW, WSR_E_UNSUPPORTED,
              "wgrad_tc: the table must loop over the whole output");
  WSR_REQUIRE(d->up == 1 || (d->up == 2 && t->in_sub == 1 && t->GH == 2 * d->H && t->GW == 2 * d->W), WSR_E_UNSUPPORTED, "wgrad_tc: bad upsample table");
  WSR_REQUIRE(t->in_sub == 1 || (d->H % 2 == 0 && d->W % 2 == 0), WSR_E_UNSUPPORTED, "wgrad_tc: in_sub 2 needs even H, W");
  WSR_REQUIRE(d->dbias == nullptr, WSR_E_UNSUPPORTED, "wgrad_tc: bias gradient is a separate call (wsr_col_sums)");

  WgParams p;
  memset(&p, 0, sizeof(p));
  // loop grid (pixels of one K segment): the dY grid, or the low-resolution grid per output phase for upsample
  const int LW = d->up == 2 ? d->W : t->GW, LH = d->up == 2 ? d->H : t->GH;
  p.t1 = LW < 64 ? LW : 64;
  p.t2 = 64 / p.t1; if (p.t2 > LH) p.t2 = LH;
  p.t3 = 64 / (p.t1 * p.t2); if (p.t3 > d->N) p.t3 = d->N;
  WSR_REQUIRE(p.t1 * p.t2 * p.t3 == 64, WSR_E_UNSUPPORTED, "wgrad_tc: cannot form 64-pixel tiles from N=%d H=%d W=%d", d->N, LH, LW);
  p.g1 = wg_cdiv(LW, p.t1); p.g2 = wg_cdiv(LH, p.t2); p.g3 = wg_cdiv(d->N, p.t3);
  p.Cout = d->Cout; p.Cin = d->Cin;
  p.dw = d->dw; p.stap = d->dw_stap; p.sco = d->dw_sco; p.sci = d->dw_sci;
  p.ntaps = t->ntaps;
  const uint32_t box[4] = {64, (uint32_t)p.t1, (uint32_t)p.t2, (uint32_t)p.t3};

  // ---- dY views
  if (d->up == 1) {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)t->OW, (uint64_t)t->OH, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)d->dy_ld * 2, (uint64_t)t->OW * d->dy_ld * 2, (uint64_t)t->OH * t->OW * d->dy_ld * 2};
    rc = encode_map(&p.amap[0], d->dy, 4, dims, str, box);
    if (rc) return rc;
    for (int i = 1; i < 4; ++i) p.amap[i] = p.amap[0];
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
        uint64_t str[3] = {(uint64_t)d->dy_ld * 4, (uint64_t)t->OW * d->dy_ld * 4, (uint64_t)t->OH * t->OW * d->dy_ld * 2};
        const __nv_bfloat16* base = (const __nv_bfloat16*)d->dy + ((long long)py * t->OW + px) * d->dy_ld;
        rc = encode_map(&p.amap[py * 2 + px], base, 4, dims, str, box);
        if (rc) return rc;
      }
  }
  // ---- X views
  if (t->in_sub == 1) {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)d->x_ld * 2, (uint64_t)d->W * d->x_ld * 2, (uint64_t)d->H * d->W * d->x_ld * 2};
    rc = encode_map(&p.bmap[0], d->x, 4, dims, str, box);
    if (rc) return rc;
    for (int i = 1; i < 4; ++i) p.bmap[i] = p.bmap[0];
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W / 2, (uint64_t)d->H / 2, (uint64_t)d->N};
        uint64_t str[3] = {(uint64_t)d->x_ld * 4, (uint64_t)d->W * d->x_ld * 4, (uint64_t)d->H * d->W * d->x_ld * 2};
        const __nv_bfloat16* base = (const __nv_bfloat16*)d->x + ((long long)py * d->W + px) * d->x_ld;
        rc = encode_map(&p.bmap[py * 2 + px], base, 4, dims, str, box);
        if (rc) return rc;
      }
  }
  // ---- K segments per tap
  p.nseg = d->up == 2 ? 4 : 1;
  for (int i = 0; i < t->ntaps; ++i) {
    p.wtap[i] = t->wtap[i];
    if (d->up == 2) {
      // output pixel (2u + py, 2v + px) reads x[floor((2u + py + dy) / 2)] = x[u + floor((py + dy) / 2)]
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          WgSeg& s = p.seg[i][py * 2 + px];
          s.amap = (int16_t)(py * 2 + px); s.bmap = 0;
          s.dy = (int16_t)floordiv2(py + t->dy[i]); s.dx = (int16_t)floordiv2(px + t->dx[i]);
        }
    } else {
      WgSeg& s = p.seg[i][0];
      s.amap = 0; s.bmap = (int16_t)(t->in_sub == 2 ? t->py[i] * 2 + t->px[i] : 0);
      s.dy = (int16_t)t->dy[i]; s.dx = (int16_t)t->dx[i];
    }
  }
  const int bn = d->Cin > 128 ? 256 : (d->Cin > 64 ? 128 : 64);
  p.m_tiles = wg_cdiv(d->Cout, 128);
  p.n_tiles = wg_cdiv(d->Cin, bn);
  const int tiles = p.ntaps * p.m_tiles * p.n_tiles;
  const int total_kb = p.nseg * p.g1 * p.g2 * p.g3;
  // K splits: one CTA per SM is resident (the operand ring takes most of the shared memory), so the launch runs in
  // ceil(tiles * splits / SMs) waves of ceil(total_kb / splits) K blocks each, plus a fixed per-CTA cost (prologue, TMEM round
  // trip, the red.global.add epilogue) of roughly a dozen K blocks.  The former rule ceil(SMs / tiles) overshot one wave by a few
  // CTAs for most layers (9 taps x 17 splits = 153 CTAs on 148 SMs: a second, almost empty wave doubled the time).
  int splits = 1;
  {
    const int sms = sm_count();
    double best = 1e30;
    const int smax = total_kb < 8 * sms ? total_kb : 8 * sms;
    for (int sp = 1; sp <= smax; ++sp) {
      const int waves = wg_cdiv(tiles * sp, sms);
      const double cost = (double)waves * ((double)wg_cdiv(total_kb, sp) + 12.0);
      if (cost < best) { best = cost; splits = sp; }
      if (tiles * sp > 4 * sms && wg_cdiv(total_kb, sp) < 12) break;
    }
  }
  p.splits = splits;
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 256: return launch_wgrad<256>(p, st);
    case 128: return launch_wgrad<128>(p, st);
    default: return launch_wgrad<64>(p, st);
  }
}
