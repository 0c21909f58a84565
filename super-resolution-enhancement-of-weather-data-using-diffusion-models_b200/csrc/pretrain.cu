// pretrain.cu -- the prior pre-training loss of the reference (models/simple_cnn/loss.py:9-80) and the ReLU
// backward mask of the SimpleCNN prior (models/simple_cnn/Simple_CNN.py:24-32).
//
//   image_compare_loss(x, y) = alpha * fft_mse_loss + beta * dwt_mse_loss
//     fft_mse_loss = MSE(Re fft2_ortho(x), Re fft2_ortho(y)) + MSE(Im ..., Im ...)
//                  = (1 / N) sum |F(x - y)|^2 = MSE(x, y)           (the orthonormal FFT is linear and unitary: Parseval)
//     dwt_mse_loss = sum over 4 Haar levels and the 3 detail bands of mean(band(x - y)^2)   (the DWT is linear)
//
// so the whole loss is a function of d = x - y, and because a 4-level Haar transform only mixes pixels inside aligned 16 x 16
// blocks, one CTA per block computes the loss contribution AND its gradient in shared memory: analysis down to the 1 x 1
// approximation, scaled detail coefficients (2 * beta / N_j), orthonormal synthesis back (the transform's adjoint is its inverse),
// plus 2 * alpha / N * d.  One read of x and y, one write of the gradient: HBM-bound, no FFT and no transform tensors.
#include "common.cuh"

namespace wsr {

// block b of level j (edge n = 16 >> j): thread t < n*n owns output (ty, tx); in = previous level's approximation (edge 2n)
__global__ void __launch_bounds__(256) image_compare_loss_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                                 float alpha, float beta, double inv_n0, double* __restrict__ loss,
                                                                 float* __restrict__ grad) {
  __shared__ float ll[5][16][17];          // ll[0] = d, ll[j] = approximation after level j (edge 16 >> j)
  __shared__ float band[4][3][8][9];       // detail coefficients of level j + 1 (edge 8 >> j)
  __shared__ float red[8];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t plane = blockIdx.z;
  const int gy = blockIdx.y * 16 + ty, gx = blockIdx.x * 16 + tx;
  const int64_t idx = (plane * H + gy) * (int64_t)W + gx;
  const float d = x[idx] - y[idx];
  ll[0][ty][tx] = d;
  float part = alpha * d * d * (float)inv_n0;         // alpha * MSE term
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = 8 >> j;
    if (ty < n && tx < n) {
      const float a = ll[j][2 * ty][2 * tx], b = ll[j][2 * ty][2 * tx + 1], c = ll[j][2 * ty + 1][2 * tx], e = ll[j][2 * ty + 1][2 * tx + 1];
      ll[j + 1][ty][tx] = (a + b + c + e) * 0.5f;
      const float b0 = (a + b - c - e) * 0.5f, b1 = (a - b + c - e) * 0.5f, b2 = (a - b - c + e) * 0.5f;
      // mean over planes * (H >> (j+1)) * (W >> (j+1)) coefficients = inv_n0 * 4^(j+1)
      const float wj = beta * (float)(inv_n0 * (double)(1 << (2 * (j + 1))));
      part += wj * (b0 * b0 + b1 * b1 + b2 * b2);
      band[j][0][ty][tx] = 2.f * wj * b0;            // d loss / d coefficient
      band[j][1][ty][tx] = 2.f * wj * b1;
      band[j][2][ty][tx] = 2.f * wj * b2;
    }
    __syncthreads();
  }
  // synthesis of the gradient: the approximation at the coarsest level carries no loss
  if (threadIdx.x == 0) ll[4][0][0] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 3; j >= 0; --j) {
    const int n = 8 >> j;
    if (ty < n && tx < n) {
      const float gl = ll[j + 1][ty][tx], g0 = band[j][0][ty][tx], g1 = band[j][1][ty][tx], g2 = band[j][2][ty][tx];
      ll[j][2 * ty][2 * tx] = (gl + g0 + g1 + g2) * 0.5f;
      ll[j][2 * ty][2 * tx + 1] = (gl + g0 - g1 - g2) * 0.5f;
      ll[j][2 * ty + 1][2 * tx] = (gl - g0 + g1 - g2) * 0.5f;
      ll[j][2 * ty + 1][2 * tx + 1] = (gl - g0 - g1 + g2) * 0.5f;
    }
    __syncthreads();
  }
  if (grad) grad[idx] = ll[0][ty][tx] + 2.f * alpha * (float)inv_n0 * d;
  // block reduction of the loss contribution
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int k = 0; k < 8; ++k) v += (double)red[k];
    atomicAdd(loss, v);
  }
}

// dy *= (y > 0): backward of an in-place ReLU whose OUTPUT y was kept
__global__ void relu_mask_kernel(const float* __restrict__ y, float* __restrict__ dy, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(y[i] > 0.f)) dy[i] = 0.f;
}

// dy[r][c] *= (y[r][c] > 0 ? 1 : slope) over an NHWC channel slice (rows x C, row pitches y_ld / dy_ld): backward of LeakyReLU from the
// kept OUTPUT (the sign of a leaky ReLU's output is the sign of its input)
template <typename TY, typename TD>
__global__ void lrelu_mask_kernel(const TY* __restrict__ y, int y_ld, TD* __restrict__ dy, int dy_ld, int64_t rows, int C, float slope) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int64_t r = i / C;
  const int c = (int)(i - r * C);
  if (!(ldf<TY>(y + r * y_ld + c) > 0.f)) {
    TD* p = dy + r * dy_ld + c;
    stf<TD>(p, ldf<TD>(p) * slope);
  }
}

}  // namespace wsr

using namespace wsr;

extern "C" int wsr_lrelu_mask(const void* y, int y_dtype, int y_ld, void* dy, int dy_dtype, int dy_ld, int64_t rows, int C, float slope,
                              void* stream) {
  WSR_REQUIRE(y && dy && rows > 0 && C > 0 && y_ld >= C && dy_ld >= C && valid_dtype(y_dtype) && valid_dtype(dy_dtype), WSR_E_INVALID,
              "lrelu_mask: bad argument");
  const unsigned blocks = (unsigned)((rows * C + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (y_dtype == WSR_BF16 && dy_dtype == WSR_BF16) lrelu_mask_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)y, y_ld, (__nv_bfloat16*)dy, dy_ld, rows, C, slope);
  else if (y_dtype == WSR_F32 && dy_dtype == WSR_F32) lrelu_mask_kernel<<<blocks, 256, 0, st>>>((const float*)y, y_ld, (float*)dy, dy_ld, rows, C, slope);
  else if (y_dtype == WSR_BF16) lrelu_mask_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)y, y_ld, (float*)dy, dy_ld, rows, C, slope);
  else lrelu_mask_kernel<<<blocks, 256, 0, st>>>((const float*)y, y_ld, (__nv_bfloat16*)dy, dy_ld, rows, C, slope);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_image_compare_loss(const float* x, const float* y, int planes, int H, int W, float alpha, float beta, double* loss,
                                      float* grad, void* stream) {
  WSR_REQUIRE(x && y && loss && planes > 0 && H > 0 && W > 0, WSR_E_INVALID, "image_compare_loss: bad argument");
  WSR_REQUIRE(H % 16 == 0 && W % 16 == 0, WSR_E_UNSUPPORTED, "image_compare_loss: %dx%d is not a multiple of 16 (4 Haar levels)", H, W);
  WSR_REQUIRE(planes <= 65535 && H / 16 <= 65535, WSR_E_UNSUPPORTED, "image_compare_loss: grid too large");
  const double inv_n0 = 1.0 / ((double)planes * H * W);
  image_compare_loss_kernel<<<dim3(W / 16, H / 16, planes), 256, 0, (cudaStream_t)stream>>>(x, y, H, W, alpha, beta, inv_n0, loss, grad);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_relu_mask(const float* y, float* dy, int64_t n, void* stream) {
  WSR_REQUIRE(y && dy && n > 0, WSR_E_INVALID, "relu_mask: bad argument");
  relu_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, dy, n);
  WSR_LAUNCH_OK();
  return WSR_OK;
}
