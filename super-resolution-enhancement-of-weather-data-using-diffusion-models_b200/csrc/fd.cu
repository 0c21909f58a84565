// fd.cu -- FD_Info_Spliter (resdiff/fd_info_spliter.py:37-117) and the Haar query images (resdiff/unet.py:124-132).
//
// The condition-only branch (4-D FFT, sigma, Gaussian high-pass, LF / HF feature maps) depends only on the condition,
// so it is computed ONCE per batch here instead of once per reverse step.  The transform is a direct
// separable DFT in double precision (one pass per axis, 2*(B+C+H+W) complex MACs per element): the whole batch costs a
// few milliseconds once per 1000 steps, and it reproduces the reference's transform over ALL FOUR axes (B, C, H, W).
#include <math_constants.h>

#include "common.cuh"

namespace wsr {

// twiddle table tw[j] = exp(-2*pi*i*j/L), double
__global__ void twiddle_kernel(double2* tw, int L) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L) return;
  double s, c;
  sincospi(-2.0 * (double)j / (double)L, &s, &c);
  tw[j] = make_double2(c, s);
}

// One DFT pass along an axis of length L and element stride S:  out[o, k, i] = scale * sum_n in[o, n, i] * tw[(k*n)%L]^(sign)
// index space: total = outer * L * S, element e -> (o, k, i)
__global__ void dft_axis_kernel(const float2* __restrict__ in, float2* __restrict__ out, const double2* __restrict__ tw,
                                int L, int64_t S, int64_t total, int inverse, double scale) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int64_t i = e % S;
  int64_t r = e / S;
  int k = (int)(r % L);
  int64_t o = r / L;
  const float2* src = in + o * L * S + i;
  double re = 0.0, im = 0.0;
  int idx = 0;
  for (int n = 0; n < L; ++n) {
    double2 w = tw[idx];
    if (inverse) w.y = -w.y;
    float2 v = src[(int64_t)n * S];
    re += (double)v.x * w.x - (double)v.y * w.y;
    im += (double)v.x * w.y + (double)v.y * w.x;
    idx += k;
    if (idx >= L) idx -= L;
  }
  out[e] = make_float2((float)(re * scale), (float)(im * scale));
}

__global__ void real_to_complex_kernel(const float* __restrict__ x, float2* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float2(x[i], 0.f);
}

// per (b, c): mean of Re and Im over HW, optionally weighted by the high-pass H_b -> means[(b*2C) + c], [(b*2C) + C + c]
__global__ void __launch_bounds__(256) spec_mean_kernel(const float2* __restrict__ spec, int C, int H, int W,
                                                        const float* __restrict__ sigma, double* __restrict__ means) {
  __shared__ double sre[256], sim[256];
  const int bc = blockIdx.x;
  const int b = bc / C, c = bc % C;
  const float2* p = spec + (int64_t)bc * H * W;
  double re = 0.0, im = 0.0;
  float inv2s2 = 0.f;
  if (sigma) { float s = sigma[b]; inv2s2 = 1.f / (2.f * s * s); }
  for (int i = threadIdx.x; i < H * W; i += 256) {
    float2 v = p[i];
    float hp = 1.f;
    if (sigma) {
      float u = (float)(i / W) - 0.5f * H, vv = (float)(i % W) - 0.5f * W;
      hp = 1.f - expf(-(u * u + vv * vv) * inv2s2);
    }
    re += (double)(v.x * hp);
    im += (double)(v.y * hp);
  }
  sre[threadIdx.x] = re; sim[threadIdx.x] = im;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sre[threadIdx.x] += sre[threadIdx.x + s]; sim[threadIdx.x] += sim[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    means[(int64_t)b * 2 * C + c] = sre[0] / (double)(H * W);
    means[(int64_t)b * 2 * C + C + c] = sim[0] / (double)(H * W);
  }
}

// ResSE squeeze-excite vector: se[b][ch] = sigmoid(fc2 * relu(fc0 * mean[b]))  (fd_info_spliter.py:135-146), ch < C2
__device__ void res_se_vec(const double* mean, int C2, const float* fc0, const float* fc2, int hidden, float* se) {
  float hid[64];
  for (int h = 0; h < hidden; ++h) {
    float a = 0.f;
    for (int c = 0; c < C2; ++c) a = fmaf(fc0[h * C2 + c], (float)mean[c], a);
    hid[h] = a > 0.f ? a : 0.f;
  }
  for (int c = 0; c < C2; ++c) {
    float a = 0.f;
    for (int h = 0; h < hidden; ++h) a = fmaf(fc2[c * hidden + h], hid[h], a);
    se[c] = 1.f / (1.f + expf(-a));
  }
}

// sigma[b] = min(| mean_ch( mean_hw(x_fd) * (1 + se) ) | + l/2, l - 10)       (fd_info_spliter.py:70-73)
__global__ void fd_sigma_kernel(const double* __restrict__ means, int B, int C, const float* fc0, const float* fc2,
                                int H, int W, float* __restrict__ sigma) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int C2 = 2 * C, hidden = C2 / 2;
  float se[64];
  res_se_vec(means + (int64_t)b * C2, C2, fc0, fc2, hidden, se);
  float acc = 0.f;
  for (int c = 0; c < C2; ++c) acc += (float)means[(int64_t)b * C2 + c] * (1.f + se[c]);
  float l = (float)min(H, W);
  sigma[b] = fminf(fabsf(acc / (float)C2) + 0.5f * l, l - 10.f);
}

__global__ void fd_se2_kernel(const double* __restrict__ means, int B, int C, const float* fc0, const float* fc2,
                              float* __restrict__ se2) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int C2 = 2 * C;
  float se[64];
  res_se_vec(means + (int64_t)b * C2, C2, fc0, fc2, C2 / 2, se);
  for (int c = 0; c < C2; ++c) se2[(int64_t)b * C2 + c] = se[c];
}

// apply the high-pass (spec -> out) and emit lf = cond * (ct_b + sum_ch ct_w[o][ch] * F_ch * (1 + se2[ch]))
__global__ void fd_filter_lf_kernel(const float2* __restrict__ spec, float2* __restrict__ out, const float* __restrict__ cond,
                                    const float* __restrict__ sigma, const float* __restrict__ se2, const float* __restrict__ ct_w,
                                    const float* __restrict__ ct_b, int B, int C, int H, int W, float* __restrict__ lf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*H*W
  int64_t HW = (int64_t)H * W;
  if (i >= (int64_t)B * HW) return;
  int b = (int)(i / HW);
  int p = (int)(i - (int64_t)b * HW);
  float s = sigma[b];
  float u = (float)(p / W) - 0.5f * H, v = (float)(p % W) - 0.5f * W;
  float hp = 1.f - expf(-(u * u + v * v) / (2.f * s * s));
  float fre[32], fim[32];
  for (int c = 0; c < C; ++c) {
    float2 q = spec[((int64_t)b * C + c) * HW + p];
    q.x *= hp; q.y *= hp;
    out[((int64_t)b * C + c) * HW + p] = q;
    fre[c] = q.x * (1.f + se2[(int64_t)b * 2 * C + c]);
    fim[c] = q.y * (1.f + se2[(int64_t)b * 2 * C + C + c]);
  }
  for (int o = 0; o < C; ++o) {
    float a = ct_b[o];
    for (int c = 0; c < C; ++c) a = fmaf(ct_w[o * 2 * C + c], fre[c], a);
    for (int c = 0; c < C; ++c) a = fmaf(ct_w[o * 2 * C + C + c], fim[c], a);
    lf[((int64_t)b * C + o) * HW + p] = cond[((int64_t)b * C + o) * HW + p] * a;
  }
}

__global__ void complex_abs_kernel(const float2* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float2 v = in[i]; out[i] = sqrtf(v.x * v.x + v.y * v.y); }
}

// ------------------------------------------------------------------------------------------------------------------
// backward of the condition-only branch w.r.t. sigma_resSE, HF_guided_resSE and channel_transform (the condition itself
// is data: resdiff_diffusion.py:121, so the spectrum F0 is a constant and only H(sigma), the two squeeze-excite vectors
// and the 1x1 channel transform carry gradient)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_atomic_add(float v, float* dst, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(dst, s);
  }
}

// ResSE fc backward for one sample: given d_se (gradient w.r.t. the sigmoid output), accumulates dfc0 / dfc2 and returns
// the gradient w.r.t. the squeezed mean (or skips it when d_mean == nullptr)
__device__ void res_se_bwd(const double* mean, int C2, const float* fc0, const float* fc2, int hidden, const float* d_se,
                           float* dfc0, float* dfc2, float* d_mean) {
  float a1[64], hid[64], da1[64];
  for (int h = 0; h < hidden; ++h) {
    float a = 0.f;
    for (int c = 0; c < C2; ++c) a = fmaf(fc0[h * C2 + c], (float)mean[c], a);
    a1[h] = a; hid[h] = a > 0.f ? a : 0.f; da1[h] = 0.f;
  }
  for (int c = 0; c < C2; ++c) {
    float a = 0.f;
    for (int h = 0; h < hidden; ++h) a = fmaf(fc2[c * hidden + h], hid[h], a);
    const float se = 1.f / (1.f + expf(-a));
    const float da2 = d_se[c] * se * (1.f - se);
    for (int h = 0; h < hidden; ++h) { atomicAdd(dfc2 + c * hidden + h, da2 * hid[h]); da1[h] = fmaf(da2, fc2[c * hidden + h], da1[h]); }
  }
  for (int h = 0; h < hidden; ++h) if (!(a1[h] > 0.f)) da1[h] = 0.f;
  for (int c = 0; c < C2; ++c) {
    float dm = 0.f;
    for (int h = 0; h < hidden; ++h) { atomicAdd(dfc0 + h * C2 + c, da1[h] * (float)mean[c]); dm = fmaf(da1[h], fc0[h * C2 + c], dm); }
    if (d_mean) d_mean[c] = dm;
  }
}

// grid (ceil(HW/128), B), 128 threads.  Direct part of dL/dF from the lf path, plus d ct_w, d ct_b and d se_h.
__global__ void __launch_bounds__(128) fd_bwd_lf_kernel(const float2* __restrict__ F0, const float* __restrict__ cond,
                                                        const float* __restrict__ sigma, const float* __restrict__ se2,
                                                        const float* __restrict__ ct_w, const float* __restrict__ g_lf, int C, int H,
                                                        int W, float2* __restrict__ dF, float* d_ctw, float* d_ctb, float* d_seh) {
  __shared__ float red[4];
  const int b = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int p = blockIdx.x * 128 + threadIdx.x;
  const bool ok = p < HW;
  float fre[32], fim[32], dct[32];
  float hp = 0.f;
  if (ok) {
    const float s = sigma[b];
    const float u = (float)(p / W) - 0.5f * H, v = (float)(p % W) - 0.5f * W;
    hp = 1.f - expf(-(u * u + v * v) / (2.f * s * s));
  }
  for (int c = 0; c < C; ++c) {
    float2 q = ok ? F0[((int64_t)b * C + c) * HW + p] : make_float2(0.f, 0.f);
    fre[c] = q.x * hp; fim[c] = q.y * hp;
    dct[c] = ok ? g_lf[((int64_t)b * C + c) * HW + p] * cond[((int64_t)b * C + c) * HW + p] : 0.f;
  }
  for (int o = 0; o < C; ++o) {
    block_atomic_add(dct[o], d_ctb + o, red);
    for (int c = 0; c < C; ++c) {
      block_atomic_add(dct[o] * fre[c] * (1.f + se2[(int64_t)b * 2 * C + c]), d_ctw + o * 2 * C + c, red);
      block_atomic_add(dct[o] * fim[c] * (1.f + se2[(int64_t)b * 2 * C + C + c]), d_ctw + o * 2 * C + C + c, red);
    }
  }
  for (int c = 0; c < C; ++c) {
    float dre = 0.f, dim = 0.f;
    for (int o = 0; o < C; ++o) { dre = fmaf(ct_w[o * 2 * C + c], dct[o], dre); dim = fmaf(ct_w[o * 2 * C + C + c], dct[o], dim); }
    block_atomic_add(dre * fre[c], d_seh + (int64_t)b * 2 * C + c, red);
    block_atomic_add(dim * fim[c], d_seh + (int64_t)b * 2 * C + C + c, red);
    if (ok) dF[((int64_t)b * C + c) * HW + p] = make_float2(dre * (1.f + se2[(int64_t)b * 2 * C + c]), dim * (1.f + se2[(int64_t)b * 2 * C + C + c]));
  }
}

__global__ void fd_bwd_se_kernel(const double* __restrict__ meansF, int B, int C, const float* fc0, const float* fc2,
                                 const float* __restrict__ d_seh, float* dfc0, float* dfc2, float* __restrict__ d_meanF) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int C2 = 2 * C;
  res_se_bwd(meansF + (int64_t)b * C2, C2, fc0, fc2, C2 / 2, d_seh + (int64_t)b * C2, dfc0, dfc2, d_meanF + (int64_t)b * C2);
}

// gz = g_hf * z / |z|  (gradient of |z| in torch's complex convention)
__global__ void fd_bwd_abs_kernel(const float2* __restrict__ z, const float* __restrict__ g_hf, float2* __restrict__ gz, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2 v = z[i];
  float a = sqrtf(v.x * v.x + v.y * v.y);
  float g = a > 0.f ? g_hf[i] / a : 0.f;
  gz[i] = make_float2(v.x * g, v.y * g);
}

// one block per sample: d sigma_b = sum_{c,p} (dReF * ReF0 + dImF * ImF0) * dH/dsigma, then the sigma ResSE backward
__global__ void __launch_bounds__(256) fd_bwd_sigma_kernel(const float2* __restrict__ F0, const float2* __restrict__ dF,
                                                          const float2* __restrict__ G, const float* __restrict__ d_meanF,
                                                          const float* __restrict__ sigma, const double* __restrict__ means0, int C,
                                                          int H, int W, const float* fc0, const float* fc2, float* dfc0, float* dfc2) {
  __shared__ double red[256];
  const int b = blockIdx.x;
  const int64_t HW = (int64_t)H * W;
  const float s = sigma[b];
  const float inv2s2 = 1.f / (2.f * s * s), invs3 = 1.f / (s * s * s);
  double acc = 0.0;
  for (int64_t p = threadIdx.x; p < HW; p += 256) {
    const float u = (float)(p / W) - 0.5f * H, v = (float)(p % W) - 0.5f * W;
    const float D2 = u * u + v * v;
    const float dHds = -expf(-D2 * inv2s2) * D2 * invs3;
    float dH = 0.f;
    for (int c = 0; c < C; ++c) {
      const int64_t i = ((int64_t)b * C + c) * HW + p;
      const float2 f = F0[i], d = dF[i], g = G[i];
      const float dre = d.x + g.x + d_meanF[(int64_t)b * 2 * C + c] / (float)HW;
      const float dim = d.y + g.y + d_meanF[(int64_t)b * 2 * C + C + c] / (float)HW;
      dH = fmaf(dre, f.x, dH);
      dH = fmaf(dim, f.y, dH);
    }
    acc += (double)(dH * dHds);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float dsig = (float)red[0];
    const int C2 = 2 * C;
    const double* mean = means0 + (int64_t)b * C2;
    float se[64];
    res_se_vec(mean, C2, fc0, fc2, C2 / 2, se);
    float a = 0.f;
    for (int c = 0; c < C2; ++c) a += (float)mean[c] * (1.f + se[c]);
    a /= (float)C2;
    const float l = (float)min(H, W);
    if (fabsf(a) + 0.5f * l < l - 10.f) {
      const float dacc = dsig * (a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f));
      float d_se[64];
      for (int c = 0; c < C2; ++c) d_se[c] = dacc * (float)mean[c] / (float)C2;
      res_se_bwd(mean, C2, fc0, fc2, C2 / 2, d_se, dfc0, dfc2, nullptr);
    }
  }
}

// per-step gate
__global__ void fd_gate_kernel(const float* __restrict__ ne_rows, int ne_ld, const int* __restrict__ row_index, int C, int W,
                               const float* __restrict__ fc0, const float* __restrict__ fc2, int hidden, float* __restrict__ gate) {
  __shared__ float red[8];
  __shared__ float se[32];
  const int b = blockIdx.x;
  const int r = row_index ? *row_index : b;
  const float* ne = ne_rows + (int64_t)r * ne_ld;
  float acc = 0.f;
  for (int w = threadIdx.x; w < W; w += blockDim.x) acc += ne[w];
  // block reduce (blockDim = 256)
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m += red[i];
    m /= (float)W;
    float hid[32];
    for (int h = 0; h < hidden; ++h) {
      float a = 0.f;
      for (int c = 0; c < C; ++c) a = fmaf(fc0[h * C + c], m, a);
      hid[h] = a > 0.f ? a : 0.f;
    }
    for (int c = 0; c < C; ++c) {
      float a = 0.f;
      for (int h = 0; h < hidden; ++h) a = fmaf(fc2[c * hidden + h], hid[h], a);
      se[c] = 1.f / (1.f + expf(-a));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
    int c = i / W, w = i - c * W;
    gate[((int64_t)b * C + c) * W + w] = ne[w] * (1.f + se[c]);
  }
}

__device__ __forceinline__ float stem_value(const float* x, const float* cond, const float* gate, const float* lf, const float* hf,
                                           int C, int W, int64_t HW, int64_t b, int64_t p, int ch) {
  if (ch >= 5 * C) return 0.f;
  const int k = ch / C, c = ch - k * C;
  const int64_t src = (b * C + c) * HW + p;
  switch (k) {
    case 0: return x[src];
    case 1: return cond[src];
    case 2: return x[src] * gate[(b * C + c) * W + (p % W)];
    case 3: return lf[src];
    default: return hf[src];
  }
}

// one thread per pixel; writes the first `nwrite` channels (a multiple of 8 for bf16) -- the remaining pad channels
// of the buffer were zeroed at allocation and are never touched again
template <typename T>
__global__ void stem_assemble_kernel(const float* __restrict__ x, const float* __restrict__ cond, const float* __restrict__ gate,
                                     const float* __restrict__ lf, const float* __restrict__ hf, int C, int H, int W,
                                     T* __restrict__ y, int Cpad, int nwrite, int64_t npix) {
  int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int64_t HW = (int64_t)H * W;
  const int64_t b = pix / HW, p = pix - b * HW;
  T* dst = y + pix * Cpad;
  if constexpr (sizeof(T) == 2) {
    for (int c0 = 0; c0 < nwrite; c0 += 8) {
      uint4 u;
      __nv_bfloat162* h = (__nv_bfloat162*)&u;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        h[k] = __floats2bfloat162_rn(stem_value(x, cond, gate, lf, hf, C, W, HW, b, p, c0 + 2 * k),
                                     stem_value(x, cond, gate, lf, hf, C, W, HW, b, p, c0 + 2 * k + 1));
      *(uint4*)(dst + c0) = u;
    }
  } else {
    for (int c = 0; c < nwrite; ++c) dst[c] = stem_value(x, cond, gate, lf, hf, C, W, HW, b, p, c);
  }
}

__global__ void haar_level_kernel(const float* __restrict__ ll, int h, int w, float* __restrict__ detail, float* __restrict__ ll_next, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over BC * h/2 * w/2
  if (i >= total) return;
  int w2 = w / 2, h2 = h / 2;
  int x = (int)(i % w2);
  int64_t r = i / w2;
  int y = (int)(r % h2);
  int64_t bc = r / h2;
  const float* p = ll + (bc * h + 2 * y) * w + 2 * x;
  float a = p[0], b = p[1], c = p[w], d = p[w + 1];
  detail[i] = (3.f * a - b - c - d) * 0.5f;
  ll_next[i] = (a + b + c + d) * 0.5f;
}

// PhyDiff keeps the three detail bands apart (phydiff/unet.py:274-276): bands[b][k*C + c] = band k (LH, HL, HH) of channel c
__global__ void haar_level_bands_kernel(const float* __restrict__ ll, int C, int h, int w, float* __restrict__ bands,
                                        float* __restrict__ ll_next, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*C * h/2 * w/2
  if (i >= total) return;
  int w2 = w / 2, h2 = h / 2;
  int x = (int)(i % w2);
  int64_t r = i / w2;
  int y = (int)(r % h2);
  int64_t bc = r / h2;
  int64_t b = bc / C; int c = (int)(bc - b * C);
  const float* p = ll + (bc * h + 2 * y) * w + 2 * x;
  float a = p[0], bb = p[1], cc = p[w], d = p[w + 1];
  const int64_t plane = (int64_t)h2 * w2;
  float* o = bands + (b * 3 * C + c) * plane + (int64_t)y * w2 + x;
  o[0] = (a + bb - cc - d) * 0.5f;
  o[(int64_t)C * plane] = (a - bb + cc - d) * 0.5f;
  o[(int64_t)2 * C * plane] = (a - bb - cc + d) * 0.5f;
  ll_next[i] = (a + bb + cc + d) * 0.5f;
}

// PhyDiff stencil channels (phydiff/unet.py:189-196,311-314): x / y forward differences and the 5-point Laplacian of the
// REFLECT-padded condition, each summed over the image channels.  out: (B, 3, H, W) fp32.
__global__ void phy_stencils_kernel(const float* __restrict__ cond, int C, int H, int W, float* __restrict__ out, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B * H * W
  if (i >= total) return;
  int x = (int)(i % W);
  int64_t r = i / W;
  int y = (int)(r % H);
  int64_t b = r / H;
  const int xl = x == 0 ? 1 : x - 1, xr = x == W - 1 ? W - 2 : x + 1;
  const int yu = y == 0 ? 1 : y - 1, yd = y == H - 1 ? H - 2 : y + 1;
  float dx = 0.f, dy = 0.f, lap = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* p = cond + (b * C + c) * (int64_t)H * W;
    const float v = p[(int64_t)y * W + x], vr = p[(int64_t)y * W + xr], vl = p[(int64_t)y * W + xl];
    const float vd = p[(int64_t)yd * W + x], vu = p[(int64_t)yu * W + x];
    dx += vr - v;
    dy += vd - v;
    lap += (vu + vl - 4.f * v) + (vr + vd);        // same tap order as the 3x3 cross-correlation (row-major taps)
  }
  const int64_t plane = (int64_t)H * W;
  float* o = out + b * 3 * plane + (int64_t)y * W + x;
  o[0] = dx; o[plane] = dy; o[2 * plane] = lap;
}

}  // namespace wsr

using namespace wsr;
static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }
static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

namespace wsr {
// workspace of wsr_fd_precompute; the first four complex buffers and the small vectors are read back by wsr_fd_backward
struct FdWork {
  float2 *buf0, *buf1, *F0, *Z;
  double2* tw;
  double *means0, *meansF;
  float *sigma, *se2;
};
static int64_t fd_work_layout(void* work, int B, int C, int H, int W, FdWork* w) {
  const int64_t n = (int64_t)B * C * H * W;
  int64_t maxL = B > C ? B : C; if (H > maxL) maxL = H; if (W > maxL) maxL = W;
  char* wp = (char*)work;
  char* w0 = wp;
  FdWork t;
  t.buf0 = (float2*)wp; wp += align256(n * 8);
  t.buf1 = (float2*)wp; wp += align256(n * 8);
  t.F0 = (float2*)wp; wp += align256(n * 8);
  t.Z = (float2*)wp; wp += align256(n * 8);
  t.tw = (double2*)wp; wp += align256(maxL * 16);
  t.means0 = (double*)wp; wp += align256((int64_t)B * 2 * C * 8);
  t.meansF = (double*)wp; wp += align256((int64_t)B * 2 * C * 8);
  t.sigma = (float*)wp; wp += align256((int64_t)B * 4);
  t.se2 = (float*)wp; wp += align256((int64_t)B * 2 * C * 4);
  if (w) *w = t;
  return (int64_t)(wp - w0) + 1024;
}

// 4-D DFT over (B, C, H, W): `cur` holds the input, `nxt` is scratch; returns the buffer that holds the result
static float2* dft4(float2* cur, float2* nxt, double2* tw, int B, int C, int H, int W, int inverse, bool scale, cudaStream_t st) {
  const int64_t n = (int64_t)B * C * H * W;
  const int L[4] = {W, H, C, B};
  const int64_t S[4] = {1, W, (int64_t)H * W, (int64_t)C * H * W};
  for (int ax = 0; ax < 4; ++ax) {
    if (L[ax] == 1) continue;
    twiddle_kernel<<<nblk(L[ax], 128), 128, 0, st>>>(tw, L[ax]);
    dft_axis_kernel<<<nblk(n, 128), 128, 0, st>>>(cur, nxt, tw, L[ax], S[ax], n, inverse, scale ? 1.0 / (double)L[ax] : 1.0);
    float2* t = cur; cur = nxt; nxt = t;
  }
  return cur;
}
}  // namespace wsr

extern "C" int64_t wsr_fd_precompute_workspace_bytes(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return -1;
  return fd_work_layout(nullptr, B, C, H, W, nullptr);
}

extern "C" int wsr_fd_precompute(const float* cond, int B, int C, int H, int W, const float* sigma_fc0,
                                 const float* sigma_fc2, const float* hf_fc0, const float* hf_fc2, const float* ct_w,
                                 const float* ct_b, int out_ch, float* lf, float* hf, void* work, void* stream) {
  WSR_REQUIRE(cond && sigma_fc0 && sigma_fc2 && hf_fc0 && hf_fc2 && ct_w && ct_b && lf && hf && work, WSR_E_INVALID, "fd_precompute: null pointer");
  WSR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, WSR_E_INVALID, "fd_precompute: bad shape");
  WSR_REQUIRE(out_ch == C, WSR_E_UNSUPPORTED, "fd_precompute: out_channel (%d) must equal image channels (%d)", out_ch, C);
  WSR_REQUIRE(C <= 32, WSR_E_UNSUPPORTED, "fd_precompute: C=%d > 32", C);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)B * C * H * W;
  FdWork w;
  fd_work_layout(work, B, C, H, W, &w);

  real_to_complex_kernel<<<nblk(n, 256), 256, 0, st>>>(cond, w.buf0, n);
  WSR_LAUNCH_OK();
  float2* spec = dft4(w.buf0, w.buf1, w.tw, B, C, H, W, 0, false, st);
  WSR_LAUNCH_OK();
  WSR_CUDA_OK(cudaMemcpyAsync(w.F0, spec, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  // sigma from the unfiltered spectrum
  spec_mean_kernel<<<B * C, 256, 0, st>>>(w.F0, C, H, W, nullptr, w.means0);
  fd_sigma_kernel<<<nblk(B, 64), 64, 0, st>>>(w.means0, B, C, sigma_fc0, sigma_fc2, H, W, w.sigma);
  // squeeze-excite of the FILTERED spectrum
  spec_mean_kernel<<<B * C, 256, 0, st>>>(w.F0, C, H, W, w.sigma, w.meansF);
  fd_se2_kernel<<<nblk(B, 64), 64, 0, st>>>(w.meansF, B, C, hf_fc0, hf_fc2, w.se2);
  fd_filter_lf_kernel<<<nblk((int64_t)B * H * W, 128), 128, 0, st>>>(w.F0, w.buf0, cond, w.sigma, w.se2, ct_w, ct_b, B, C, H, W, lf);
  WSR_LAUNCH_OK();
  // inverse transform over all four axes
  float2* z = dft4(w.buf0, w.buf1, w.tw, B, C, H, W, 1, true, st);
  WSR_LAUNCH_OK();
  WSR_CUDA_OK(cudaMemcpyAsync(w.Z, z, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  complex_abs_kernel<<<nblk(n, 256), 256, 0, st>>>(w.Z, hf, n);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int64_t wsr_fd_backward_workspace_bytes(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return -1;
  const int64_t n = (int64_t)B * C * H * W;
  int64_t maxL = B > C ? B : C; if (H > maxL) maxL = H; if (W > maxL) maxL = W;
  return 3 * align256(n * 8) + align256(maxL * 16) + 2 * align256((int64_t)B * 2 * C * 4) + 1024;
}

extern "C" int wsr_fd_backward(const float* cond, int B, int C, int H, int W, const float* sigma_fc0, const float* sigma_fc2,
                               const float* hf_fc0, const float* hf_fc2, const float* ct_w, const float* g_lf, const float* g_hf,
                               const void* work, void* bwork, float* d_sigma_fc0, float* d_sigma_fc2, float* d_hf_fc0,
                               float* d_hf_fc2, float* d_ct_w, float* d_ct_b, void* stream) {
  WSR_REQUIRE(cond && sigma_fc0 && sigma_fc2 && hf_fc0 && hf_fc2 && ct_w && g_lf && g_hf && work && bwork && d_sigma_fc0 && d_sigma_fc2 &&
                  d_hf_fc0 && d_hf_fc2 && d_ct_w && d_ct_b, WSR_E_INVALID, "fd_backward: null pointer");
  WSR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && C <= 32, WSR_E_INVALID, "fd_backward: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)B * C * H * W;
  FdWork w;
  fd_work_layout(const_cast<void*>(work), B, C, H, W, &w);
  int64_t maxL = B > C ? B : C; if (H > maxL) maxL = H; if (W > maxL) maxL = W;
  char* bp = (char*)bwork;
  float2* dF = (float2*)bp; bp += align256(n * 8);
  float2* g0 = (float2*)bp; bp += align256(n * 8);
  float2* g1 = (float2*)bp; bp += align256(n * 8);
  double2* tw = (double2*)bp; bp += align256(maxL * 16);
  float* d_seh = (float*)bp; bp += align256((int64_t)B * 2 * C * 4);
  float* d_meanF = (float*)bp;
  WSR_CUDA_OK(cudaMemsetAsync(d_seh, 0, (size_t)B * 2 * C * 4, st));
  dim3 grid((unsigned)(((int64_t)H * W + 127) / 128), B);
  fd_bwd_lf_kernel<<<grid, 128, 0, st>>>(w.F0, cond, w.sigma, w.se2, ct_w, g_lf, C, H, W, dF, d_ct_w, d_ct_b, d_seh);
  WSR_LAUNCH_OK();
  fd_bwd_se_kernel<<<nblk(B, 64), 64, 0, st>>>(w.meansF, B, C, hf_fc0, hf_fc2, d_seh, d_hf_fc0, d_hf_fc2, d_meanF);
  fd_bwd_abs_kernel<<<nblk(n, 256), 256, 0, st>>>(w.Z, g_hf, g0, n);
  WSR_LAUNCH_OK();
  // gradient of ifftn: (1/N) * fftn of the cotangent
  float2* G = dft4(g0, g1, tw, B, C, H, W, 0, true, st);
  WSR_LAUNCH_OK();
  fd_bwd_sigma_kernel<<<B, 256, 0, st>>>(w.F0, dF, G, d_meanF, w.sigma, w.means0, C, H, W, sigma_fc0, sigma_fc2, d_sigma_fc0, d_sigma_fc2);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_fd_gate(const float* ne_rows, int ne_ld, const int* row_index, int B, int C, int W, const float* fc0,
                           const float* fc2, int hidden, float* gate, void* stream) {
  WSR_REQUIRE(ne_rows && fc0 && fc2 && gate && B > 0 && C > 0 && W > 0 && ne_ld >= W, WSR_E_INVALID, "fd_gate: bad argument");
  WSR_REQUIRE(C <= 32 && hidden > 0 && hidden <= 32, WSR_E_UNSUPPORTED, "fd_gate: C or hidden > 32");
  fd_gate_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(ne_rows, ne_ld, row_index, C, W, fc0, fc2, hidden, gate);
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_stem_assemble(const float* x, const float* cond, const float* gate, const float* lf, const float* hf,
                                 int B, int C, int H, int W, void* y, int y_dtype, int Cpad, void* stream) {
  WSR_REQUIRE(x && cond && gate && lf && hf && y && valid_dtype(y_dtype), WSR_E_INVALID, "stem_assemble: null pointer");
  WSR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && Cpad >= 5 * C, WSR_E_INVALID, "stem_assemble: bad shape");
  const int64_t npix = (int64_t)B * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (y_dtype == WSR_BF16) {
    int nwrite = (5 * C + 7) / 8 * 8;
    WSR_REQUIRE(Cpad % 8 == 0 && nwrite <= Cpad && (((uintptr_t)y) & 15) == 0, WSR_E_UNSUPPORTED, "stem_assemble: bf16 needs Cpad %% 8 == 0");
    stem_assemble_kernel<__nv_bfloat16><<<nblk(npix, 128), 128, 0, st>>>(x, cond, gate, lf, hf, C, H, W, (__nv_bfloat16*)y, Cpad, nwrite, npix);
  } else {
    stem_assemble_kernel<float><<<nblk(npix, 128), 128, 0, st>>>(x, cond, gate, lf, hf, C, H, W, (float*)y, Cpad, 5 * C, npix);
  }
  WSR_LAUNCH_OK();
  return WSR_OK;
}

extern "C" int wsr_haar_detail_sums(const float* img, int B, int C, int H, int W, int levels, float* out, float* ll_work, void* stream) {
  WSR_REQUIRE(img && out && ll_work && B > 0 && C > 0 && levels > 0, WSR_E_INVALID, "haar: bad argument");
  WSR_REQUIRE((H % (1 << levels)) == 0 && (W % (1 << levels)) == 0, WSR_E_UNSUPPORTED, "haar: H, W must be divisible by 2^levels");
  cudaStream_t st = (cudaStream_t)stream;
  const float* ll = img;
  int h = H, w = W;
  float* lw0 = ll_work;
  float* lw1 = ll_work + (int64_t)B * C * (H / 2) * (W / 2);
  float* o = out;
  for (int j = 0; j < levels; ++j) {
    int64_t total = (int64_t)B * C * (h / 2) * (w / 2);
    float* nxt = (j % 2 == 0) ? lw0 : lw1;
    haar_level_kernel<<<nblk(total, 256), 256, 0, st>>>(ll, h, w, o, nxt, total);
    WSR_LAUNCH_OK();
    o += total;
    ll = nxt; h /= 2; w /= 2;
  }
  return WSR_OK;
}

extern "C" int wsr_haar_detail_bands(const float* img, int B, int C, int H, int W, int levels, float* out, float* ll_work, void* stream) {
  WSR_REQUIRE(img && out && ll_work && B > 0 && C > 0 && levels > 0, WSR_E_INVALID, "haar bands: bad argument");
  WSR_REQUIRE((H % (1 << levels)) == 0 && (W % (1 << levels)) == 0, WSR_E_UNSUPPORTED, "haar bands: H, W must be divisible by 2^levels");
  cudaStream_t st = (cudaStream_t)stream;
  const float* ll = img;
  int h = H, w = W;
  float* lw0 = ll_work;
  float* lw1 = ll_work + (int64_t)B * C * (H / 2) * (W / 2);
  float* o = out;
  for (int j = 0; j < levels; ++j) {
    int64_t total = (int64_t)B * C * (h / 2) * (w / 2);
    float* nxt = (j % 2 == 0) ? lw0 : lw1;
    haar_level_bands_kernel<<<nblk(total, 256), 256, 0, st>>>(ll, C, h, w, o, nxt, total);
    WSR_LAUNCH_OK();
    o += 3 * total;
    ll = nxt; h /= 2; w /= 2;
  }
  return WSR_OK;
}

extern "C" int wsr_phy_stencils(const float* cond, int B, int C, int H, int W, float* out, void* stream) {
  WSR_REQUIRE(cond && out && B > 0 && C > 0 && H >= 2 && W >= 2, WSR_E_INVALID, "phy_stencils: bad argument (reflect padding needs H, W >= 2)");
  int64_t total = (int64_t)B * H * W;
  phy_stencils_kernel<<<nblk(total, 256), 256, 0, (cudaStream_t)stream>>>(cond, C, H, W, out, total);
  WSR_LAUNCH_OK();
  return WSR_OK;
}
