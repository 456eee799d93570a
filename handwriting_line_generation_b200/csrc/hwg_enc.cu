// Memory-bound passes of the perceptual encoder Encoder2 (reference model/autoencoder.py:341-410) and of the perceptual
// loss the trainer builds on it (trainer/hw_with_style_trainer.py:740-748) that the discriminator's passes (hwg_disc.cu)
// do not already cover: the residual addition with the GroupNorm statistics of the sum, and the L1 loss between the two
// halves of a feature tensor with its gradient.  NHWC bf16, 16-byte vectors of 8 channels, fp32 arithmetic; every kernel's
// roofline is HBM (operands read once, result written once).
// GPU parity: tests/test_enc_gpu.py.
#include "common.cuh"

namespace hwg {
namespace {

constexpr int ET = 256;      // threads per block

__device__ __forceinline__ void unpack8e(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8e(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// ---- y = a + b; stats[n,c] += (sum y, sum y^2) of the STORED (bf16-rounded) values -----------------------------
// grid (blocks per sample, N); thread = (pixel lane, 8-channel vector).  The per-channel block reduction goes through a
// shared scratch tile (no shared-memory atomics), one global atomic per channel per block.
template <int CV, bool STATS>
__global__ void __launch_bounds__(ET)
add_stats_kernel(const uint4* a, const uint4* b, uint4* y, long long HW,   // y may alias a or b: no __restrict__
                 float* __restrict__ stats) {
  constexpr int LANES = ET / CV;
  __shared__ float red[STATS ? LANES : 1][2 * CV * 8 + 1];
  const int cv = threadIdx.x % CV, lane = threadIdx.x / CV;
  const long long base = (long long)blockIdx.y * HW * CV;
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  for (long long p = (long long)blockIdx.x * LANES + lane; p < HW; p += (long long)gridDim.x * LANES) {
    const long long i = base + p * CV + cv;
    float fa[8], fb[8];
    unpack8e(a[i], fa);
    unpack8e(b[i], fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) fa[k] += fb[k];
    const uint4 o = pack8e(fa);
    y[i] = o;
    if (STATS) {
      unpack8e(o, fb);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += fb[k]; q[k] = fmaf(fb[k], fb[k], q[k]); }
    }
  }
  if (STATS) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { red[lane][2 * (cv * 8 + k)] = s[k]; red[lane][2 * (cv * 8 + k) + 1] = q[k]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CV * 8; i += ET) {
      float t = 0.f;
#pragma unroll 4
      for (int l = 0; l < LANES; ++l) t += red[l][i];
      atomicAdd(stats + (long long)blockIdx.y * (2 * CV * 8) + i, t);     // stats [N][C][2]: index = 2*c + {0,1}
    }
  }
}

// ---- L1 between the halves of f = [orig ; recon]: loss += loss_scale * sum |r - o|, g = grad_scale * sign(r - o) ----
template <bool F32>
__global__ void __launch_bounds__(ET)
l1_halves_kernel(const void* __restrict__ f, long long half_vec, float loss_scale, float grad_scale,
                 float* __restrict__ loss, uint4* __restrict__ g) {
  __shared__ float red[ET / 32];
  float acc = 0.f;
  for (long long v = (long long)blockIdx.x * ET + threadIdx.x; v < half_vec; v += (long long)gridDim.x * ET) {
    float o[8], r[8];
    if (F32) {
      const float4* p = reinterpret_cast<const float4*>(f);
      const float4 o0 = p[2 * v], o1 = p[2 * v + 1], r0 = p[2 * (half_vec + v)], r1 = p[2 * (half_vec + v) + 1];
      o[0] = o0.x; o[1] = o0.y; o[2] = o0.z; o[3] = o0.w; o[4] = o1.x; o[5] = o1.y; o[6] = o1.z; o[7] = o1.w;
      r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
    } else {
      const uint4* p = reinterpret_cast<const uint4*>(f);
      unpack8e(p[v], o);
      unpack8e(p[half_vec + v], r);
    }
    float gs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = r[k] - o[k];
      acc += fabsf(d);
      gs[k] = d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f);        // torch's sign(): 0 at 0
    }
    if (g) g[v] = pack8e(gs);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < ET / 32; ++i) t += red[i];
    atomicAdd(loss, t * loss_scale);
  }
}

}  // namespace
}  // namespace hwg

using namespace hwg;

extern "C" int hwg_add_stats(const void* a, const void* b, void* y, int N, int64_t HW, int C, float* stats,
                             void* stream) {
  HWG_REQUIRE(a && b && y && N > 0 && HW > 0, "hwg_add_stats: bad argument");
  HWG_REQUIRE(C == 16 || C == 32 || C == 64 || C == 128 || C == 256, "hwg_add_stats: C=%d must be 16, 32, 64, 128 or 256", C);
  const int CV = C / 8, LANES = ET / CV;
  long long bx = (HW + (long long)LANES * 8 - 1) / ((long long)LANES * 8);       // >= 8 pixels per lane
  const long long cap = (148LL * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const dim3 grid((unsigned)bx, (unsigned)N);
  cudaStream_t s = (cudaStream_t)stream;
  const uint4 *ap = reinterpret_cast<const uint4*>(a), *bp = reinterpret_cast<const uint4*>(b);
  uint4* yp = reinterpret_cast<uint4*>(y);
#define ADD_CASE(CVV)                                                                                  \
  case CVV:                                                                                            \
    if (stats) add_stats_kernel<CVV, true><<<grid, ET, 0, s>>>(ap, bp, yp, HW, stats);                 \
    else add_stats_kernel<CVV, false><<<grid, ET, 0, s>>>(ap, bp, yp, HW, stats);                      \
    break;
  switch (CV) {
    ADD_CASE(2) ADD_CASE(4) ADD_CASE(8) ADD_CASE(16) ADD_CASE(32)
    default: break;
  }
#undef ADD_CASE
  return check_launch("add_stats_kernel");
}

extern "C" int hwg_l1_halves(const void* f, int dtype, int64_t half_numel, float loss_scale, float grad_scale,
                             float* loss, void* g, void* stream) {
  HWG_REQUIRE(f && loss && half_numel > 0 && half_numel % 8 == 0, "hwg_l1_halves: half_numel=%lld must be a positive multiple of 8",
              (long long)half_numel);
  HWG_REQUIRE(dtype == HWG_DT_BF16 || dtype == HWG_DT_F32, "hwg_l1_halves: bad dtype");
  HWG_REQUIRE((reinterpret_cast<uintptr_t>(f) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
              (dtype == HWG_DT_F32 || half_numel % 8 == 0), "hwg_l1_halves: 16-byte alignment");
  const long long hv = half_numel / 8;
  long long bx = (hv + ET - 1) / ET;
  if (bx > 148LL * 8) bx = 148LL * 8;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HWG_DT_F32)
    l1_halves_kernel<true><<<(unsigned)bx, ET, 0, s>>>(f, hv, loss_scale, grad_scale, loss, reinterpret_cast<uint4*>(g));
  else
    l1_halves_kernel<false><<<(unsigned)bx, ET, 0, s>>>(f, hv, loss_scale, grad_scale, loss, reinterpret_cast<uint4*>(g));
  return check_launch("l1_halves_kernel");
}
