// edge.cu -- the steps either side of the sampling loop, on the device: the bicubic x-scale collate that
// produces the condition 'SR' (data/dataset_builder.py:374-380), the StandardScaling transform and its inverse back to
// physical units (data/transforms.py:391-409, applied per sample and variable by _inverse_tensor, :116-138), and the
// accumulators of the validation metrics MAE / MSE / RMSE / MR (training/metrics.py:75-201).  All HBM-bound, one pass each.
#include "common.cuh"

namespace wsr {

// torch.nn.functional.interpolate(mode="bicubic", align_corners=False): cubic convolution with A = -0.75, source index
// (o + 0.5) / scale - 0.5, taps clamped to the image (aten/src/ATen/native/UpSample.h: cubic_convolution1/2, upsample_get_value_bounded)
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__global__ void bicubic_upsample_kernel(const float* __restrict__ src, int h, int w, int H, int W, float rscale_h, float rscale_w,
                                        float* __restrict__ dst, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over planes * H * W
  if (i >= total) return;
  const int ox = (int)(i % W);
  int64_t r = i / W;
  const int oy = (int)(r % H);
  const int64_t plane = r / H;
  const float sy = (oy + 0.5f) * rscale_h - 0.5f, sx = (ox + 0.5f) * rscale_w - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  const int iy = (int)fy, ix = (int)fx;
  float cy[4], cx[4];
  cubic_coeffs(sy - fy, cy);
  cubic_coeffs(sx - fx, cx);
  const float* p = src + plane * (int64_t)h * w;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = min(max(iy - 1 + a, 0), h - 1);
    float row = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int xx = min(max(ix - 1 + b, 0), w - 1);
      row = fmaf(p[(int64_t)yy * w + xx], cx[b], row);
    }
    acc = fmaf(row, cy[a], acc);
  }
  dst[i] = acc;
}

// Fast path for the power-of-two scales the reference uses (4 everywhere, 8 for the stress configuration): one thread per SOURCE
// pixel produces its S x S block of outputs.  For these scales the fractional offsets (o + 0.5) / S - 0.5 - floor(.) take only S
// values (exact in fp32), so the S x 4 tap weights are computed once per block into shared memory; a thread loads the 5 x 5 clamped
// source window once (L1 hits: a 32x64 plane is 8 KB), runs the horizontal pass for the 5 rows and all S phases, then the vertical
// pass per output row, and stores S rows of S contiguous floats (16-byte stores, a warp writes 32 * S * 4 contiguous bytes per row).
// The accumulation order (horizontal fmaf chain over b = 0..3, then vertical chain over a = 0..3) is the generic kernel's, so the
// results are bit-identical to it; instructions per output pixel drop from ~100 to ~13 and the kernel becomes store-bound.
template <int S>
__global__ void __launch_bounds__(256) bicubic_block_kernel(const float* __restrict__ src, int h, int w, float* __restrict__ dst, int64_t n_src) {
  __shared__ float coef[S][4];
  if (threadIdx.x < S) {
    const float s = (threadIdx.x + 0.5f) * (1.f / (float)S) - 0.5f;
    float c[4];
    cubic_coeffs(s - floorf(s), c);
#pragma unroll
    for (int b = 0; b < 4; ++b) coef[threadIdx.x][b] = c[b];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over planes * h * w
  if (i >= n_src) return;
  const int kx = (int)(i % w);
  const int64_t r = i / w;
  const int ky = (int)(r % h);
  const int64_t plane = r / h;
  const float* p = src + plane * (int64_t)h * w;
  float win[5][5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    const int yy = min(max(ky - 2 + a, 0), h - 1);
#pragma unroll
    for (int b = 0; b < 5; ++b) win[a][b] = __ldg(p + (int64_t)yy * w + min(max(kx - 2 + b, 0), w - 1));
  }
  // horizontal pass: phase jx < S/2 starts at source column kx - 2, the others at kx - 1
  float hrow[5][S];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int jx = 0; jx < S; ++jx) {
      const int o = jx < S / 2 ? 0 : 1;
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) row = fmaf(win[a][o + b], coef[jx][b], row);
      hrow[a][jx] = row;
    }
  const int W = w * S;
  float* q = dst + (plane * (int64_t)h * S + (int64_t)ky * S) * W + (int64_t)kx * S;
#pragma unroll
  for (int jy = 0; jy < S; ++jy) {
    const int o = jy < S / 2 ? 0 : 1;
    float out[S];
#pragma unroll
    for (int jx = 0; jx < S; ++jx) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) acc = fmaf(hrow[o + a][jx], coef[jy][a], acc);
      out[jx] = acc;
    }
#pragma unroll
    for (int v = 0; v < S / 4; ++v) *(float4*)(q + (int64_t)jy * W + 4 * v) = make_float4(out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
  }
}

// y = (x - mean[plane]) / std[plane]  (inverse = 0)   or   y = std[plane] * x + mean[plane]  (inverse = 1)
// vectorised variant: one plane per blockIdx.y (no per-element division by hw), 16-byte accesses, four of them in flight per thread
__global__ void __launch_bounds__(256) standard_scale_vec_kernel(const float4* __restrict__ x, int64_t hw4, const float* __restrict__ mean,
                                                                 const float* __restrict__ stdv, int inverse, float4* __restrict__ y) {
  const int64_t plane = blockIdx.y;
  const float m = mean[plane], s = stdv[plane];
  const float4* xp = x + plane * hw4;
  float4* yp = y + plane * hw4;
  for (int64_t i0 = (int64_t)blockIdx.x * 1024 + threadIdx.x; i0 < hw4; i0 += (int64_t)gridDim.x * 1024) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i0 + 256 * k < hw4) v[k] = __ldcs(xp + i0 + 256 * k);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + 256 * k >= hw4) break;
      float4 o;
      if (inverse) {
        o.x = __fadd_rn(__fmul_rn(s, v[k].x), m); o.y = __fadd_rn(__fmul_rn(s, v[k].y), m);
        o.z = __fadd_rn(__fmul_rn(s, v[k].z), m); o.w = __fadd_rn(__fmul_rn(s, v[k].w), m);
      } else {
        o.x = __fdiv_rn(v[k].x - m, s); o.y = __fdiv_rn(v[k].y - m, s); o.z = __fdiv_rn(v[k].z - m, s); o.w = __fdiv_rn(v[k].w - m, s);
      }
      __stcs(yp + i0 + 256 * k, o);
    }
  }
}

__global__ void standard_scale_kernel(const float* __restrict__ x, int64_t hw, const float* __restrict__ mean, const float* __restrict__ stdv,
                                      int inverse, float* __restrict__ y, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t plane = i / hw;
  const float m = mean[plane], s = stdv[plane];
  y[i] = inverse ? __fadd_rn(__fmul_rn(s, x[i]), m) : __fdiv_rn(x[i] - m, s);      // exact IEEE ops: --use_fast_math would approximate the division
}

// acc[0] += sum |d|, acc[1] += sum d^2, acc[2] += sum d  with d = scale[plane] * (pred - target)  (scale == nullptr: 1)
__global__ void __launch_bounds__(256) error_sums_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t hw,
                                                         const float* __restrict__ scale, double* acc, int64_t total) {
  double s_abs = 0.0, s_sq = 0.0, s_d = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float d = pred[i] - target[i];
    if (scale) d *= scale[i / hw];
    s_abs += (double)fabsf(d);
    s_sq += (double)d * (double)d;
    s_d += (double)d;
  }
  __shared__ double red[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_abs += __shfl_xor_sync(0xffffffffu, s_abs, o);
    s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o);
    s_d += __shfl_xor_sync(0xffffffffu, s_d, o);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][warp] = s_abs; red[1][warp] = s_sq; red[2][warp] = s_d; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int k = 0; k < 8; ++k) v += red[threadIdx.x][k];
    atomicAdd(acc + threadIdx.x, v);
  }
}

// vectorised variant: blockIdx.y = plane (one scale per block, no division), 16-byte loads, the eight differences of two loads are
// summed in fp32 and only those partial sums enter the double accumulators (8 elements per double operation instead of 1)
__global__ void __launch_bounds__(256) error_sums_vec_kernel(const float4* __restrict__ pred, const float4* __restrict__ target, int64_t hw4,
                                                             const float* __restrict__ scale, double* acc) {
  const int64_t plane = blockIdx.y;
  const float sc = scale ? scale[plane] : 1.f;
  const float4* pp = pred + plane * hw4;
  const float4* tp = target + plane * hw4;
  double s_abs = 0.0, s_sq = 0.0, s_d = 0.0;
  for (int64_t i0 = (int64_t)blockIdx.x * 512 + threadIdx.x; i0 < hw4; i0 += (int64_t)gridDim.x * 512) {
    float4 a[2], b[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const bool ok = i0 + 256 * k < hw4;
      a[k] = ok ? __ldcs(pp + i0 + 256 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
      b[k] = ok ? __ldcs(tp + i0 + 256 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float fa = 0.f, fq = 0.f, fd = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float d0 = (a[k].x - b[k].x) * sc, d1 = (a[k].y - b[k].y) * sc, d2 = (a[k].z - b[k].z) * sc, d3 = (a[k].w - b[k].w) * sc;
      fa += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
      fq += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
      fd += (d0 + d1) + (d2 + d3);
    }
    s_abs += (double)fa; s_sq += (double)fq; s_d += (double)fd;
  }
  __shared__ double red[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_abs += __shfl_xor_sync(0xffffffffu, s_abs, o);
    s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o);
    s_d += __shfl_xor_sync(0xffffffffu, s_d, o);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][warp] = s_abs; red[1][warp] = s_sq; red[2][warp] = s_d; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int k = 0; k < 8; ++k) v += red[threadIdx.x][k];
    atomicAdd(acc + threadIdx.x, v);
  }
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace wsr

using namespace wsr;

extern "C" int wsr_bicubic_upsample(const float* src, int planes, int h, int w, int scale, float* dst, void* stream) {
  WSR_REQUIRE(src && dst && planes > 0 && h > 0 && w > 0 && scale >= 1, WSR_E_INVALID, "bicubic_upsample: bad argument");
  const int H = h * scale, W = w * scale;
  const int64_t total = (int64_t)planes * H * W;
  // torch uses the reciprocal of the USER scale factor when one is given (area_pixel_compute_scale with scale_factor)
  const float rs = 1.f / (float)scale;
  const int64_t n_src = (int64_t)planes * h * w;
  if ((scale == 4 || scale == 8) && aligned16(dst)) {
    if (scale == 4) bicubic_block_kernel<4><<<(unsigned)((n_src + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, h, w, dst, n_src);
    else bicubic_block_kernel<8><<<(unsigned)((n_src + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, h, w, dst, n_src);
    WSR_LAUNCH_OK();
    return WSR_OK;
  }
  bicubic_upsample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, h, w, H, W, rs, rs, dst, total);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_standard_scale(const float* x, int planes, int64_t hw, const float* mean, const float* stdv, int inverse, float* y,
                                  void* stream) {
  WSR_REQUIRE(x && y && mean && stdv && planes > 0 && hw > 0, WSR_E_INVALID, "standard_scale: bad argument");
  const int64_t total = (int64_t)planes * hw;
  if (hw % 4 == 0 && aligned16(x) && aligned16(y) && planes <= 65535) {
    const int64_t hw4 = hw / 4;
    int64_t bx = (hw4 + 1023) / 1024;
    const int64_t cap = (148 * 16 + planes - 1) / planes;            // ~16 blocks per SM over the whole grid
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    standard_scale_vec_kernel<<<dim3((unsigned)bx, (unsigned)planes), 256, 0, (cudaStream_t)stream>>>((const float4*)x, hw4, mean, stdv,
                                                                                                    inverse ? 1 : 0, (float4*)y);
    WSR_LAUNCH_OK();
    return WSR_OK;
  }
  standard_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, hw, mean, stdv, inverse ? 1 : 0, y, total);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_error_sums(const float* pred, const float* target, int planes, int64_t hw, const float* scale, double* acc, void* stream) {
  WSR_REQUIRE(pred && target && acc && planes > 0 && hw > 0, WSR_E_INVALID, "error_sums: bad argument");
  const int64_t total = (int64_t)planes * hw;
  if (hw % 4 == 0 && aligned16(pred) && aligned16(target) && planes <= 65535) {
    // without a scale the planes are indistinguishable: re-cut the flat range into up to 64 pseudo-planes to fill the grid
    int64_t pl = planes, hw4 = hw / 4;
    if (scale == nullptr) {
      const int64_t n4 = total / 4;
      pl = 1;
      while (pl < 64 && (n4 % (pl * 2)) == 0 && n4 / (pl * 2) >= 512) pl *= 2;
      hw4 = n4 / pl;
    }
    int64_t bx = (hw4 + 511) / 512;
    const int64_t cap = (148 * 8 + pl - 1) / pl;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    error_sums_vec_kernel<<<dim3((unsigned)bx, (unsigned)pl), 256, 0, (cudaStream_t)stream>>>((const float4*)pred, (const float4*)target, hw4, scale, acc);
    WSR_LAUNCH_OK();
    return WSR_OK;
  }
  int64_t blocks = (total + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  error_sums_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, hw, scale, acc, total);
  WSR_LAUNCH_OK();
  return WSR_OK;
}
