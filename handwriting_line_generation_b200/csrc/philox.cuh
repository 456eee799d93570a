// Counter-based N(0,1) noise for NoiseInjection (reference pure_gen.py:206,212 draws
// torch.randn_like): Philox4x32-10 keyed by (seed), counter (element/4, subsequence), then
// Box-Muller.  Any element can be generated independently, so the noise is fused into the
// producing kernel's epilogue instead of being materialised in HBM.
#pragma once
#include <stdint.h>

namespace hwg {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

// four independent standard normals for elements 4*idx4 .. 4*idx4+3
__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t subseq, uint64_t idx4) {
  uint4 r = philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)subseq,
                                     (uint32_t)(subseq >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k24 = 1.0f / 16777216.0f;
  float u0 = (float)((r.x >> 8) + 1u) * k24, u1 = (float)(r.y >> 8) * k24;
  float u2 = (float)((r.z >> 8) + 1u) * k24, u3 = (float)(r.w >> 8) * k24;
  float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u1, &sa, &ca);
  __sincosf(6.283185307179586f * u3, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

__device__ __forceinline__ float normal1(uint64_t seed, uint64_t subseq, uint64_t idx) {
  float4 z = normal4(seed, subseq, idx >> 2);
  const int k = (int)(idx & 3);
  return k == 0 ? z.x : (k == 1 ? z.y : (k == 2 ? z.z : z.w));
}

}  // namespace hwg
