// philox.cuh -- counter-based Philox4x32-10 shared by the sampler noise (elementwise.cu) and the training-mode
// dropout masks (elementwise.cu forward, backward.cu backward): the backward pass regenerates the mask of the forward
// pass from (seed, tag, element index) instead of storing it.
#pragma once
#include <stdint.h>

namespace wsr {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// Dropout keep-scale of VEC consecutive elements starting at logical element index e0 (e0 % VEC == 0 for VEC in {4, 8}):
// m[i] = 0 (dropped, probability p) or 1/(1-p) (kept).  Element e uses word (e & 3) of Philox block (e >> 2).
template <int VEC>
__device__ __forceinline__ void dropout_scale(uint64_t seed, uint32_t tag, uint64_t e0, float p, float (&m)[VEC]) {
  const uint32_t thresh = (uint32_t)fminf(p * 4294967296.f, 4294967040.f);
  const float keep = 1.f / (1.f - p);
  if constexpr (VEC == 1) {
    uint64_t g = e0 >> 2;
    uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), tag, 0x44524f50u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    m[0] = c[e0 & 3] >= thresh ? keep : 0.f;
  } else {
#pragma unroll
    for (int q = 0; q < VEC / 4; ++q) {
      uint64_t g = (e0 >> 2) + q;
      uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), tag, 0x44524f50u};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
      for (int j = 0; j < 4; ++j) m[q * 4 + j] = c[j] >= thresh ? keep : 0.f;
    }
  }
}

}  // namespace wsr
