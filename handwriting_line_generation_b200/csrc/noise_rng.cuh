// Counter-based N(0,1) noise for NoiseInjection (reference pure_gen.py:206,212 draws torch.randn_like).
// The noise is generated inside the producing kernel's epilogue instead of being materialised in HBM
// (a randn tensor costs 8 bytes of traffic per element next to the 2-byte bf16 activation).  On the
// 16/32-channel layers the epilogue is instruction-bound, so the generator is built for few
// instructions per sample: one 32-bit integer hash of (element pair index, seed, subsequence), stretched to 2x32 bits, feeds
// one Box-Muller transform with 24-bit uniforms and the fast lg2/sin/cos units -> two normals for
// ~27 instructions.  Any element can be generated independently (no state), so results do not depend
// on the tiling.
#pragma once
#include <stdint.h>

namespace hwg {

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

// key = per-launch constant mixed from (seed, subsequence); computed once per thread
__device__ __forceinline__ uint2 noise_key(unsigned long long seed, unsigned long long subseq) {
  const uint32_t a = fmix32((uint32_t)seed ^ 0x9E3779B9u) ^ fmix32((uint32_t)(subseq) + 0x7F4A7C15u);
  const uint32_t b = fmix32((uint32_t)(seed >> 32) + 0x94D049BBu) ^ fmix32((uint32_t)(subseq >> 32) ^ 0xBF58476Du);
  return make_uint2(a, b);
}

// two independent standard normals for elements (2*pair, 2*pair+1)
// Round 2: the passes that draw the noise are bound by instruction issue (ncu: 74-78 % issue activity, ~13 instructions per
// element for the noise alone), so the generator was trimmed: the second 32-bit word is ONE odd multiply + xorshift of the
// (already avalanched) first word instead of a second two-round hash, and the logarithm is lg2.approx.ftz (u1 >= 2^-25 is
// never subnormal, so the compiler's denormal guard around __log2f was dead weight).
// (lo, him) = low word of the pair index and its high word times 0x9E3779B1 (a per-item constant for callers that draw
// several consecutive pairs)
__device__ __forceinline__ float2 normal_pair_lh(uint2 key, uint32_t lo, uint32_t him) {
  const uint32_t x = fmix32((lo ^ key.x) + him);
  uint32_t y = (x ^ key.y) * 0x2C1B3C6Du;
  y ^= y >> 16;
  const float u1 = fmaf((float)(x >> 8), 5.9604645e-8f, 2.9802322e-8f);   // (0,1): (k + 0.5) / 2^24
  const float ang = (float)(y >> 8) * 3.7450704e-7f;                       // 2*pi*k / 2^24
  float l2, r;                                                             // sqrt(-2 ln u1), ln = lg2 * ln2
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862944f * l2));   // one MUFU op instead of the IEEE sequence
  float s, c;
  __sincosf(ang, &s, &c);
  return make_float2(r * c, r * s);
}
__device__ __forceinline__ float2 normal_pair(uint2 key, unsigned long long pair) {
  return normal_pair_lh(key, (uint32_t)pair, (uint32_t)(pair >> 32) * 0x9E3779B1u);
}
// the eight normals of pairs pair0 .. pair0+3, pair0 a multiple of 4 (one 16-byte vector of 8 channels): the +q never
// carries into the high word, so the index arithmetic per pair is one 32-bit add
__device__ __forceinline__ void normal_oct(uint2 key, unsigned long long pair0, float (&z)[8]) {
  const uint32_t lo = (uint32_t)pair0, him = (uint32_t)(pair0 >> 32) * 0x9E3779B1u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 zz = normal_pair_lh(key, lo + (uint32_t)q, him);
    z[2 * q] = zz.x; z[2 * q + 1] = zz.y;
  }
}

__device__ __forceinline__ float normal_one(uint2 key, unsigned long long idx) {
  const float2 z = normal_pair(key, idx >> 1);
  return (idx & 1ull) ? z.y : z.x;
}

}  // namespace hwg
