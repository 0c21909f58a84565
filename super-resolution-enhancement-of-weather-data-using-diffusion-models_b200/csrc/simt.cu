// simt.cu -- fp32-accumulate SIMT kernels: implicit-GEMM convolution with a tap table, batched strided GEMM and the
// k8/s4 transposed convolution.  These are the "fp32 check mode" kernels: exact fp32
// arithmetic, any channel count, used to validate the tcgen05 path and for the once-per-sample precompute.
#include "common.cuh"

namespace wsr {

constexpr int kMaxTaps = 16;

struct ConvSimtParams {
  WsrConvDesc d;
  int ntaps;
  int dy[kMaxTaps], dx[kMaxTaps], wtap[kMaxTaps];   // input offset (in the possibly upsampled grid) and weight tap
  int GH, GW;            // loop grid (per image): one GEMM row per (gy, gx)
  int in_stride;         // input coordinate = g*in_stride + d
  int UH, UW, up;        // upsampled input extent and factor
  int out_mul, out_py, out_px;   // output pixel = g*out_mul + out_p
  int OH, OW;            // output tensor extent
};

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const WsrConvDesc& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t M = (int64_t)d.N * p.GH * p.GW;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;

  // decode the GEMM row this thread loads
  const int64_t mload = m0 + lrow;
  const bool mvalid = mload < M;
  int img = 0, gy = 0, gx = 0;
  if (mvalid) {
    img = (int)(mload / ((int64_t)p.GH * p.GW));
    int r = (int)(mload - (int64_t)img * p.GH * p.GW);
    gy = r / p.GW;
    gx = r - gy * p.GW;
  }
  const int co_load = n0 + lrow;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const T* x = (const T*)d.x;
  const T* w = (const T*)d.w;
  const int nseg = d.x2 ? 2 : 1;
  for (int seg = 0; seg < nseg; ++seg) {
    const int taps = seg == 0 ? p.ntaps : 1;
    const int Cin = seg == 0 ? d.Cin : d.Cin2;
    for (int t = 0; t < taps; ++t) {
      const T* src = nullptr;
      const T* wp = nullptr;
      if (seg == 0) {
        if (mvalid) {
          int uy = gy * p.in_stride + p.dy[t], ux = gx * p.in_stride + p.dx[t];
          if (uy >= 0 && uy < p.UH && ux >= 0 && ux < p.UW)
            src = x + ((int64_t)(img * d.H + uy / p.up) * d.W + ux / p.up) * d.x_ld;
        }
        if (co_load < d.Cout) wp = w + ((int64_t)p.wtap[t] * (d.w_rows > 0 ? d.w_rows : d.Cout) + co_load) * Cin;
      } else {
        if (mvalid) {
          int oy = gy * p.out_mul + p.out_py, ox = gx * p.out_mul + p.out_px;
          src = (const T*)d.x2 + ((int64_t)(img * p.OH + oy) * p.OW + ox) * d.x2_ld;
        }
        if (co_load < d.Cout) wp = (const T*)d.w2 + (int64_t)co_load * Cin;
      }
      for (int c0 = 0; c0 < Cin; c0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + lk + j;
          As[lk + j][lrow] = (src && c < Cin) ? ldf<T>(src + c) : 0.f;
          Bs[lk + j][lrow] = (wp && c < Cin) ? ldf<T>(wp + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int im = (int)(m / ((int64_t)p.GH * p.GW));
    int r = (int)(m - (int64_t)im * p.GH * p.GW);
    int yy = (r / p.GW) * p.out_mul + p.out_py, xx = (r % p.GW) * p.out_mul + p.out_px;
    int64_t pix = ((int64_t)im * p.OH + yy) * p.OW + xx;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= d.Cout) continue;
      float v = acc[i][j];
      if (d.bias) v += d.bias[co];
      if (d.rowvec) v += d.rowvec[(int64_t)im * d.rowvec_ld + co];
      v = apply_act(v, d.act) * d.out_scale;
      if (d.res) v += d.res_scale * ld_dt(d.res, pix * d.res_ld + co, d.res_dtype);
      if (d.res2) v += d.res2_scale * ld_dt(d.res2, pix * d.res2_ld + co, d.res2_dtype);
      st_dt(d.y, pix * d.y_ld + co, d.y_dtype, v);
    }
  }
}

static int launch_conv_simt(const ConvSimtParams& p, cudaStream_t st) {
  int64_t M = (int64_t)p.d.N * p.GH * p.GW;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((p.d.Cout + 63) / 64));
  if (p.d.x_dtype == WSR_BF16) conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else conv_simt_kernel<float><<<grid, 256, 0, st>>>(p);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

int validate_conv_desc(const WsrConvDesc* d) {
  WSR_REQUIRE(d != nullptr, WSR_E_INVALID, "conv: null descriptor");
  WSR_REQUIRE(d->x && d->w && d->y, WSR_E_INVALID, "conv: null x/w/y");
  WSR_REQUIRE(valid_dtype(d->x_dtype) && valid_dtype(d->y_dtype), WSR_E_INVALID, "conv: bad dtype");
  WSR_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, WSR_E_INVALID, "conv: bad shape");
  WSR_REQUIRE(d->ksize == 1 || d->ksize == 3, WSR_E_UNSUPPORTED, "conv: ksize %d (only 1 or 3)", d->ksize);
  WSR_REQUIRE(d->stride == 1 || d->stride == 2, WSR_E_UNSUPPORTED, "conv: stride %d (only 1 or 2)", d->stride);
  WSR_REQUIRE(!(d->upsample && d->stride != 1), WSR_E_UNSUPPORTED, "conv: upsample with stride != 1");
  WSR_REQUIRE(d->x_ld >= d->Cin && d->y_ld >= d->Cout, WSR_E_INVALID, "conv: pitch smaller than channel count");
  if (d->stride == 2) WSR_REQUIRE(d->H % 2 == 0 && d->W % 2 == 0, WSR_E_UNSUPPORTED, "conv: stride 2 needs even H, W");
  if (d->x2) WSR_REQUIRE(d->w2 && d->Cin2 > 0 && d->x2_ld >= d->Cin2, WSR_E_INVALID, "conv: bad second segment");
  if (d->res) WSR_REQUIRE(valid_dtype(d->res_dtype) && d->res_ld >= d->Cout, WSR_E_INVALID, "conv: bad residual");
  if (d->res2) WSR_REQUIRE(valid_dtype(d->res2_dtype) && d->res2_ld >= d->Cout, WSR_E_INVALID, "conv: bad residual 2");
  if (d->rowvec) WSR_REQUIRE(d->rowvec_ld >= d->Cout, WSR_E_INVALID, "conv: rowvec pitch");
  if (d->gn_stats) WSR_REQUIRE(d->gn_stats_ld >= 2 * d->Cout, WSR_E_INVALID, "conv: gn_stats pitch");
  return WSR_OK;
}

}  // namespace wsr

using namespace wsr;

extern "C" int wsr_conv_simt(const WsrConvDesc* d, void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  if (d->upsample == 2) {
    // phase-merged weights: one launch per output phase, 2x2 taps on the SOURCE grid
    WSR_REQUIRE(d->ksize == 3 && d->x2 == nullptr, WSR_E_UNSUPPORTED, "conv: merged upsample needs ksize 3, no second segment");
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      ConvSimtParams q;
      q.d = *d;
      q.ntaps = 4;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          int t = a * 2 + b;
          q.dy[t] = (py == 0 ? -1 : 0) + a; q.dx[t] = (px == 0 ? -1 : 0) + b;
          q.wtap[t] = ph * 4 + t;
        }
      q.up = 1; q.UH = d->H; q.UW = d->W; q.in_stride = 1;
      q.GH = d->H; q.GW = d->W; q.OH = 2 * d->H; q.OW = 2 * d->W;
      q.out_mul = 2; q.out_py = py; q.out_px = px;
      rc = launch_conv_simt(q, (cudaStream_t)stream);
      if (rc) return rc;
    }
    if (d->gn_stats) return wsr_gn_stats(d->y, d->y_dtype, d->N, 4 * d->H * d->W, d->Cout, d->y_ld, d->gn_stats, d->gn_stats_ld, stream);
    return WSR_OK;
  }
  ConvSimtParams p;
  p.d = *d;
  const int up = d->upsample ? 2 : 1;
  const int pad = (d->ksize - 1) / 2;
  p.ntaps = d->ksize * d->ksize;
  for (int ky = 0; ky < d->ksize; ++ky)
    for (int kx = 0; kx < d->ksize; ++kx) {
      int t = ky * d->ksize + kx;
      p.dy[t] = ky - pad; p.dx[t] = kx - pad; p.wtap[t] = t;
    }
  p.up = up; p.UH = d->H * up; p.UW = d->W * up;
  p.in_stride = d->stride;
  p.OH = p.UH / d->stride; p.OW = p.UW / d->stride;
  p.GH = p.OH; p.GW = p.OW;
  p.out_mul = 1; p.out_py = 0; p.out_px = 0;
  rc = launch_conv_simt(p, (cudaStream_t)stream);
  if (rc) return rc;
  if (d->gn_stats) return wsr_gn_stats(d->y, d->y_dtype, d->N, p.OH * p.OW, d->Cout, d->y_ld, d->gn_stats, d->gn_stats_ld, stream);
  return WSR_OK;
}

namespace wsr {
int validate_taps(const WsrTapTable* t) {
  WSR_REQUIRE(t != nullptr, WSR_E_INVALID, "taps: null table");
  WSR_REQUIRE(t->ntaps > 0 && t->ntaps <= WSR_MAX_TAPS, WSR_E_INVALID, "taps: ntaps=%d", t->ntaps);
  WSR_REQUIRE(t->in_sub == 1 || t->in_sub == 2 || t->in_sub == 4, WSR_E_UNSUPPORTED, "taps: in_sub=%d (1, 2; 4 on the SIMT kernels)", t->in_sub);
  WSR_REQUIRE(t->GH > 0 && t->GW > 0 && t->OH > 0 && t->OW > 0 && t->out_mul >= 1, WSR_E_INVALID, "taps: bad grid");
  WSR_REQUIRE(t->out_py >= 0 && t->out_px >= 0 && (t->GH - 1) * t->out_mul + t->out_py < t->OH && (t->GW - 1) * t->out_mul + t->out_px < t->OW,
              WSR_E_INVALID, "taps: output grid exceeds the output tensor");
  for (int i = 0; i < t->ntaps; ++i)
    WSR_REQUIRE(t->py[i] >= 0 && t->py[i] < t->in_sub && t->px[i] >= 0 && t->px[i] < t->in_sub && t->wtap[i] >= 0, WSR_E_INVALID,
                "taps: bad tap %d", i);
  return WSR_OK;
}
}  // namespace wsr

extern "C" int wsr_conv_taps_simt(const WsrConvDesc* d, const WsrTapTable* t, void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  rc = validate_taps(t);
  if (rc) return rc;
  ConvSimtParams p;
  p.d = *d;
  p.ntaps = t->ntaps;
  for (int i = 0; i < t->ntaps; ++i) {
    p.dy[i] = t->in_sub * t->dy[i] + t->py[i];
    p.dx[i] = t->in_sub * t->dx[i] + t->px[i];
    p.wtap[i] = t->wtap[i];
  }
  p.up = 1; p.UH = d->H; p.UW = d->W; p.in_stride = t->in_sub;
  p.GH = t->GH; p.GW = t->GW; p.OH = t->OH; p.OW = t->OW;
  p.out_mul = t->out_mul; p.out_py = t->out_py; p.out_px = t->out_px;
  rc = launch_conv_simt(p, (cudaStream_t)stream);
  if (rc) return rc;
  if (d->gn_stats) return wsr_gn_stats(d->y, d->y_dtype, d->N, t->OH * t->OW, d->Cout, d->y_ld, d->gn_stats, d->gn_stats_ld, stream);
  return WSR_OK;
}

extern "C" int wsr_conv_transpose_k8s4(const void* x, int x_dtype, int N, int H, int W, int Cin, int x_ld,
                                       const void* w, const float* bias, int Cout, void* y, int y_dtype, int y_ld,
                                       void* stream) {
  WSR_REQUIRE(x && w && y, WSR_E_INVALID, "convT: null pointer");
  WSR_REQUIRE(valid_dtype(x_dtype) && valid_dtype(y_dtype), WSR_E_INVALID, "convT: bad dtype");
  WSR_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && x_ld >= Cin && y_ld >= Cout, WSR_E_INVALID, "convT: bad shape");
  // y[4i+py] gets x[i + hi] * w[ky_hi] + x[i + hi - 1] * w[ky_hi + 4], hi = (py >= 2), ky_hi = (py + 2) % 4
  for (int py = 0; py < 4; ++py)
    for (int px = 0; px < 4; ++px) {
      ConvSimtParams p;
      WsrConvDesc& d = p.d;
      d = WsrConvDesc{};
      d.x = x; d.x_dtype = x_dtype; d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.x_ld = x_ld;
      d.w = w; d.ksize = 1; d.stride = 1; d.upsample = 0; d.Cout = Cout;
      d.bias = bias; d.act = WSR_ACT_NONE; d.out_scale = 1.f;
      d.y = y; d.y_dtype = y_dtype; d.y_ld = y_ld;
      p.ntaps = 4;
      int hy = py >= 2 ? 1 : 0, hx = px >= 2 ? 1 : 0;
      int kyh = (py + 2) % 4, kxh = (px + 2) % 4;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          int t = a * 2 + b;
          p.dy[t] = hy - a; p.dx[t] = hx - b;
          p.wtap[t] = (kyh + 4 * a) * 8 + (kxh + 4 * b);
        }
      p.up = 1; p.UH = H; p.UW = W; p.in_stride = 1;
      p.GH = H; p.GW = W; p.OH = 4 * H; p.OW = 4 * W;
      p.out_mul = 4; p.out_py = py; p.out_px = px;
      int rc = launch_conv_simt(p, (cudaStream_t)stream);
      if (rc) return rc;
    }
  return WSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// batched strided GEMM
// ------------------------------------------------------------------------------------------------------------------
namespace wsr {

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const WsrGemmDesc g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int bz = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const TA* A = (const TA*)g.a + (int64_t)bz * g.a_sb;
  const TB* B = (const TB*)g.b + (int64_t)bz * g.b_sb;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + lk + j;
      int m = m0 + lrow, n = n0 + lrow;
      As[lk + j][lrow] = (m < g.M && k < g.K) ? ldf<TA>(A + (int64_t)m * g.a_sm + (int64_t)k * g.a_sk) : 0.f;
      Bs[lk + j][lrow] = (n < g.N && k < g.K) ? ldf<TB>(B + (int64_t)n * g.b_sn + (int64_t)k * g.b_sk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      if (g.bias) v += g.bias[n];
      if (g.res) v += ld_dt(g.res, (int64_t)bz * g.res_sb + (int64_t)m * g.res_sm + (int64_t)n * g.res_sn, g.res_dtype);
      st_dt(g.d, (int64_t)bz * g.d_sb + (int64_t)m * g.d_sm + (int64_t)n * g.d_sn, g.d_dtype, v);
    }
  }
}

int validate_gemm_desc(const WsrGemmDesc* g) {
  WSR_REQUIRE(g != nullptr, WSR_E_INVALID, "gemm: null descriptor");
  WSR_REQUIRE(g->a && g->b && g->d, WSR_E_INVALID, "gemm: null a/b/d");
  WSR_REQUIRE(valid_dtype(g->a_dtype) && valid_dtype(g->b_dtype) && valid_dtype(g->d_dtype), WSR_E_INVALID, "gemm: bad dtype");
  WSR_REQUIRE(g->batch > 0 && g->M > 0 && g->N > 0 && g->K > 0, WSR_E_INVALID, "gemm: bad shape");
  if (g->res) WSR_REQUIRE(valid_dtype(g->res_dtype), WSR_E_INVALID, "gemm: bad residual dtype");
  return WSR_OK;
}

}  // namespace wsr

extern "C" int wsr_gemm_simt(const WsrGemmDesc* g, void* stream) {
  int rc = validate_gemm_desc(g);
  if (rc) return rc;
  WSR_REQUIRE(g->batch <= 65535, WSR_E_UNSUPPORTED, "gemm_simt: batch > 65535");
  dim3 grid((g->M + 63) / 64, (g->N + 63) / 64, g->batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (g->a_dtype == WSR_BF16 && g->b_dtype == WSR_BF16) gemm_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(*g);
  else if (g->a_dtype == WSR_BF16) gemm_simt_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(*g);
  else if (g->b_dtype == WSR_BF16) gemm_simt_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(*g);
  else gemm_simt_kernel<float, float><<<grid, 256, 0, st>>>(*g);
  WSR_LAUNCH_OK();
  return WSR_OK;
}
