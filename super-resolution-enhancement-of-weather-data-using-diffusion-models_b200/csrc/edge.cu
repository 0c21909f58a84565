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

// y = (x - mean[plane]) / std[plane]  (inverse = 0)   or   y = std[plane] * x + mean[plane]  (inverse = 1)
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

}  // namespace wsr

using namespace wsr;

extern "C" int wsr_bicubic_upsample(const float* src, int planes, int h, int w, int scale, float* dst, void* stream) {
  WSR_REQUIRE(src && dst && planes > 0 && h > 0 && w > 0 && scale >= 1, WSR_E_INVALID, "bicubic_upsample: bad argument");
  const int H = h * scale, W = w * scale;
  const int64_t total = (int64_t)planes * H * W;
  // torch uses the reciprocal of the USER scale factor when one is given (area_pixel_compute_scale with scale_factor)
  const float rs = 1.f / (float)scale;
  bicubic_upsample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, h, w, H, W, rs, rs, dst, total);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_standard_scale(const float* x, int planes, int64_t hw, const float* mean, const float* stdv, int inverse, float* y,
                                  void* stream) {
  WSR_REQUIRE(x && y && mean && stdv && planes > 0 && hw > 0, WSR_E_INVALID, "standard_scale: bad argument");
  const int64_t total = (int64_t)planes * hw;
  standard_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, hw, mean, stdv, inverse ? 1 : 0, y, total);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_error_sums(const float* pred, const float* target, int planes, int64_t hw, const float* scale, double* acc, void* stream) {
  WSR_REQUIRE(pred && target && acc && planes > 0 && hw > 0, WSR_E_INVALID, "error_sums: bad argument");
  const int64_t total = (int64_t)planes * hw;
  int64_t blocks = (total + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  error_sums_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, hw, scale, acc, total);
  WSR_LAUNCH_OK();
  return WSR_OK;
}
