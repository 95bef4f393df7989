// Philox4x32-10 + Box-Muller: 4 unit normals per counter. Shared by the stand-alone noise kernel (elementwise.cu), the GEMM noise
// epilogue (gemm.cu) and the pooled-embedding noise (stack2_ops.cu): element i of a tensor always takes lane (i & 3) of counter
// offset + (i >> 2), so the fused forms draw exactly the normals the stand-alone kernel would draw for the same flattened tensor.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vmae {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long counter) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  // uniforms in (0,1]
  const float u0 = ((float)r.x + 1.0f) * 2.3283064365386963e-10f, u1 = (float)r.y * 2.3283064365386963e-10f;
  const float u2 = ((float)r.z + 1.0f) * 2.3283064365386963e-10f, u3 = (float)r.w * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

}  // namespace vmae
