// HBM-bound backward passes of the recognizer (sm_100a): NHWC bf16 gradients, 16-byte vectors of
// 8 channels per thread, per-channel reductions folded warp -> shared -> one global atomic per block.
#include "common.cuh"
#include <type_traits>
#include "noise_rng.cuh"
#include "stream.cuh"
#include <math_constants.h>

namespace hwg {

__device__ __forceinline__ void unpack8b(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

constexpr int BW_THREADS = 256;
constexpr int BW_ITER = 8;
constexpr int BW_U = 4;   // items whose loads are issued back to back before any is consumed (bytes in flight)

// The streaming kernels below keep their per-channel constants in SHARED memory ([k][C] floats, read as two
// float4 per use) instead of 40-60 registers per thread: at ~100 registers only two 256-thread blocks fit an SM and
// a one-load-at-a-time loop leaves ~16 KB in flight per SM (measured 1.3-2.8 TB/s); with the constants in shared
// memory and BW_U independent 16-byte loads per tensor issued up front the same kernels keep >= 64 KB in flight.
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Per-channel reduction of K x 8 per-thread partial sums.  The thread's channel vector `cv` is fixed
// (CV divides BW_THREADS).  sacc: [K][C] shared floats (zeroed by the caller, followed by a barrier).
// Threads that share a channel vector within a warp are folded with shuffles (CV < 32); the per-warp partials are
// then combined through a shared scratch tile with plain stores and a strided sum (no shared-memory atomics: with
// 256 threads adding 16 values each into <= 1024 addresses those were ~45 % of the kernel time, stall_short_sb),
// and every block issues ONE global atomic per (k, channel).
template <int K>
__device__ __forceinline__ void channel_reduce_in(float (&acc)[K][8], int cv, int CV, int C, float* red,
                                                  float* gout, int gstride) {
  constexpr int RW = K * 8 * 32 + 4;                        // floats per warp row of `red`
  if (CV < 32) {
    for (int off = 16; off >= CV; off >>= 1) {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] += __shfl_xor_sync(0xffffffffu, acc[k][j], off);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slots = CV < 32 ? CV : 32;                      // distinct channel vectors held by one warp
  if (lane < slots) {
    float4* dst = reinterpret_cast<float4*>(&red[warp * RW + lane * (K * 8)]);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      dst[2 * k] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
      dst[2 * k + 1] = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
    }
  }
  __syncthreads();
  // warp w, lane l holds channel vector (w*32 + l) % CV  (CV >= 32)  or  l % CV  (CV < 32)
  const int nwarps = BW_THREADS / 32;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    const int v = c >> 3, j = c & 7;                        // channel vector, element
    float t = 0.f;
    if (CV >= 32) {
      const int wstep = CV / 32;                            // warps w0, w0 + wstep, ... hold vector v
      for (int w = v / 32; w < nwarps; w += wstep) t += red[w * RW + (v & 31) * (K * 8) + k * 8 + j];
    } else {
      for (int w = 0; w < nwarps; ++w) t += red[w * RW + v * (K * 8) + k * 8 + j];
    }
    atomicAdd(&gout[c * gstride + k], t);
  }
}
template <int K>
constexpr int channel_reduce_scratch_floats() { return (BW_THREADS / 32) * (K * 8 * 32 + 4); }

// with a static scratch tile (kernels that have no shared-memory ring to reuse)
template <int K>
__device__ __forceinline__ void channel_reduce(float (&acc)[K][8], int cv, int CV, int C, float* /*unused*/,
                                               float* gout, int gstride) {
  __shared__ __align__(16) float red[channel_reduce_scratch_floats<K>()];
  channel_reduce_in<K>(acc, cv, CV, C, red, gout, gstride);
}

// ---- log-softmax backward ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
logsoftmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ lp, int T, int B, int C, int Cp,
                      __nv_bfloat16* __restrict__ gz, float* __restrict__ dbias) {
  extern __shared__ float sb[];  // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;  // row = t*B + b
  if (row < (long long)T * B) {
    const int t = (int)(row / B), b = (int)(row - (long long)t * B);
    const float* gr = g + row * C;
    const float* lr = lp + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += gr[c];
    s = warp_sum(s);
    __nv_bfloat16* out = gz + ((long long)b * T + t) * Cp;
    for (int c = lane; c < Cp; c += 32) {
      float v = 0.f;
      if (c < C) {
        v = gr[c] - __expf(lr[c]) * s;
        atomicAdd(&sb[c], v);
      }
      out[c] = __float2bfloat16_rn(v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&dbias[i], sb[i]);
}

// ---- BatchNorm (+ReLU) backward ----------------------------------------------------------------
// Shared memory of the bulk-staged kernels: [ring | mbarriers | per-channel constants]; the ring doubles as the
// scratch tile of the final per-channel reduction.
struct StreamSmem {
  unsigned char* ring; uint64_t* bars; float* cst;
};
template <int NS>
__device__ __forceinline__ StreamSmem stream_smem(unsigned char* raw) {
  StreamSmem m;
  m.ring = raw + ((128u - (sm100::smem_u32(raw) & 127u)) & 127u);
  m.bars = reinterpret_cast<uint64_t*>(m.ring + (size_t)ST_STAGES * NS * ST_CHUNK * 16);
  m.cst = reinterpret_cast<float*>(m.bars + ST_BAR_SLOTS);   // 16-byte aligned (ld.shared.v4)
  return m;
}
static size_t stream_kernel_smem(int nstream, int const_floats) {
  return stream_smem_bytes(nstream) + ST_BAR_SLOTS * 8 + (size_t)const_floats * sizeof(float);
}
static_assert(BW_THREADS == ST_THREADS, "channel_reduce assumes the streaming block size");
static_assert((size_t)ST_STAGES * ST_CHUNK * 16 >= sizeof(float) * (BW_THREADS / 32) * (2 * 8 * 32 + 4),
              "ring too small to double as the reduction scratch");

__global__ void __launch_bounds__(ST_THREADS)
bn_bwd_reduce_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const float* __restrict__ coef,
                     const float* __restrict__ save, long long rows, int C, int relu, float* __restrict__ sums) {
  extern __shared__ unsigned char smraw[];
  const StreamSmem sm = stream_smem<2>(smraw);      // constants a, b, mean, rstd [4][C]
  const int CV = C / 8;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sm.cst[c] = coef[2 * c]; sm.cst[C + c] = coef[2 * c + 1];
    sm.cst[2 * C + c] = save[2 * c]; sm.cst[3 * C + c] = save[2 * c + 1];
  }
  const int cv = threadIdx.x % CV;
  const float* kc = sm.cst + cv * 8;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const uint4* const src[2] = {g, z};
  stream_chunks<2>(src, rows * CV, blockIdx.x, gridDim.x, sm.ring, sm.bars, [&](long long, const uint4 (&v)[2]) {
    float gf[8], zf[8];
    unpack8b(v[0], gf);
    unpack8b(v[1], zf);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float a[4], b[4], mean[4], rstd[4];
      ld4s(kc + 4 * hf, a); ld4s(kc + C + 4 * hf, b); ld4s(kc + 2 * C + 4 * hf, mean); ld4s(kc + 3 * C + 4 * hf, rstd);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * hf + jj;
        const float gy = (!relu || fmaf(a[jj], zf[j], b[jj]) > 0.f) ? gf[j] : 0.f;
        acc[0][j] += gy;
        acc[1][j] += gy * (zf[j] - mean[jj]) * rstd[jj];
      }
    }
  });
  channel_reduce_in<2>(acc, cv, CV, C, reinterpret_cast<float*>(sm.ring), sums, 2);
}

__global__ void __launch_bounds__(ST_THREADS)
bn_bwd_apply_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const float* __restrict__ coef,
                    const float* __restrict__ save, const float* __restrict__ weight,
                    const float* __restrict__ sums, long long rows, long long norm_rows, int C, int relu,
                    uint4* __restrict__ gz, float* __restrict__ dconv_bias) {
  extern __shared__ unsigned char smraw[];
  const StreamSmem sm = stream_smem<2>(smraw);      // constants a, b, sc, P, Q [5][C]
  const int CV = C / 8;
  const float invM = 1.f / (float)norm_rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // gz = sc*(gy - k0 - xhat*k1), xhat = (z - mean)*rstd   ==   sc*gy + P*z + Q
    const float mean = save[2 * c], rstd = save[2 * c + 1];
    const float sc = (weight ? weight[c] : 1.f) * rstd;
    const float k0 = sums[2 * c] * invM, k1 = sums[2 * c + 1] * invM;
    sm.cst[c] = coef[2 * c]; sm.cst[C + c] = coef[2 * c + 1]; sm.cst[2 * C + c] = sc;
    sm.cst[3 * C + c] = -sc * k1 * rstd;
    sm.cst[4 * C + c] = sc * (k1 * rstd * mean - k0);
  }
  const int cv = threadIdx.x % CV;
  const float* kc = sm.cst + cv * 8;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  const uint4* const src[2] = {g, z};
  stream_chunks<2>(src, rows * CV, blockIdx.x, gridDim.x, sm.ring, sm.bars, [&](long long item, const uint4 (&v)[2]) {
    float gf[8], zf[8], o[8];
    unpack8b(v[0], gf);
    unpack8b(v[1], zf);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float a[4], b[4], sc[4], P[4], Q[4];
      ld4s(kc + 4 * hf, a); ld4s(kc + C + 4 * hf, b); ld4s(kc + 2 * C + 4 * hf, sc);
      ld4s(kc + 3 * C + 4 * hf, P); ld4s(kc + 4 * C + 4 * hf, Q);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * hf + jj;
        const float gy = (!relu || fmaf(a[jj], zf[j], b[jj]) > 0.f) ? gf[j] : 0.f;
        o[j] = fmaf(sc[jj], gy, fmaf(P[jj], zf[j], Q[jj]));
        acc[0][j] += o[j];
      }
    }
    gz[item] = pack8b(o);
  });
  if (dconv_bias) channel_reduce_in<1>(acc, cv, CV, C, reinterpret_cast<float*>(sm.ring), dconv_bias, 1);
}

// ---- ReLU + MaxPool backward (gather) ------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS)
relu_maxpool_bwd_kernel(const uint4* __restrict__ ga, const uint4* __restrict__ c, int N, int H, int W, int C,
                        int kh, int kw, int sh, int sw, int ph, int pw, int Ho, int Wo, uint4* __restrict__ gc,
                        float* __restrict__ dbias) {
  extern __shared__ float sacc[];  // [1][C]
  const int CV = C / 8;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cv = threadIdx.x % CV;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  const long long total = (long long)N * H * W * CV;
  const long long base = (long long)blockIdx.x * BW_THREADS * BW_ITER;
  for (int it = 0; it < BW_ITER; ++it) {
    const long long item = base + it * BW_THREADS + threadIdx.x;
    if (item >= total) break;
    long long pix = item / CV;
    const int w = (int)(pix % W); pix /= W;
    const int h = (int)(pix % H);
    const int n = (int)(pix / H);
    float me[8], o[8];
    unpack8b(c[item], me);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    // pooling windows (ho,wo) that contain (h,w): ho*sh - ph <= h < ho*sh - ph + kh
    const int ho_hi = min(Ho - 1, (h + ph) / sh);
    const int ho_lo = max(0, (h + ph - kh + sh) / sh);   // ceil((h+ph-kh+1)/sh)
    const int wo_hi = min(Wo - 1, (w + pw) / sw);
    const int wo_lo = max(0, (w + pw - kw + sw) / sw);
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        // is (h,w) the FIRST maximum of this window (row-major scan, strict > to replace)?
        bool first[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) first[j] = true;
        for (int i = 0; i < kh; ++i) {
          const int hh = ho * sh - ph + i;
          if (hh < 0 || hh >= H) continue;
          for (int j2 = 0; j2 < kw; ++j2) {
            const int ww = wo * sw - pw + j2;
            if (ww < 0 || ww >= W || (hh == h && ww == w)) continue;
            float f[8];
            unpack8b(c[(((long long)n * H + hh) * W + ww) * CV + cv], f);
            const bool before = (hh < h) || (hh == h && ww < w);
#pragma unroll
            for (int j = 0; j < 8; ++j) first[j] = first[j] && (before ? (f[j] < me[j]) : (f[j] <= me[j]));
          }
        }
        float gf[8];
        unpack8b(ga[(((long long)n * Ho + ho) * Wo + wo) * CV + cv], gf);
#pragma unroll
        for (int j = 0; j < 8; ++j) if (first[j]) o[j] += gf[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = me[j] > 0.f ? o[j] : 0.f;   // ReLU mask (c is the post-ReLU activation)
      acc[0][j] += o[j];
    }
    gc[item] = pack8b(o);
  }
  if (dbias) channel_reduce<1>(acc, cv, CV, C, sacc, dbias, 1);
}

// packed bf16x2 helpers of the window-centric pooling passes (comparisons of bf16 values are exact, so the selection logic
// runs on HMNMX2 / HSET2 / HMUL2 without unpacking: ~5x fewer instructions than the fp32 compare-and-branch chains)
using bf2 = __nv_bfloat162;
__device__ __forceinline__ bf2 as_bf2(uint32_t u) { return *reinterpret_cast<bf2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(bf2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ bf2 bf2_zero() { return as_bf2(0u); }
__device__ __forceinline__ uint32_t& u4c(uint4& u, int k) { return k == 0 ? u.x : k == 1 ? u.y : k == 2 ? u.z : u.w; }
__device__ __forceinline__ uint32_t u4c(const uint4& u, int k) { return k == 0 ? u.x : k == 1 ? u.y : k == 2 ? u.z : u.w; }

// ---- ReLU + MaxPool backward, window-centric specialisations -----------------------------------------------
// The two poolings of CNNOnlyHWR (cnn_only_hwr.py:44-56): MODE 0 = MaxPool2d(2,2) (disjoint windows) and
// MODE 1 = MaxPool2d((2,2),(2,1),(0,1)) (windows overlap by one column; padding is -inf).  A thread owns one 8-channel
// vector of a row pair (rows 2ho, 2ho+1) — MODE 0: columns 2wo, 2wo+1 (one window, four outputs); MODE 1: column w
// (windows wo = w and w+1, two outputs) — so every activation is read once per window that contains it by the thread
// that needs it, with no per-pixel division chains or data-dependent loops.  First maximum in row-major scan order
// wins, as in ATen's max_pool2d_with_indices.
// three resident blocks per SM: the pass waits on its (up to eight) 16-byte loads per item (ncu, round 2: long_scoreboard 65 %,
// 99 registers = two blocks), so resident warps are what it needs
// IDX = unsigned when the item count fits 31 bits (always, at the sizes of this path): the flat index is decoded with
// 32-bit divisions instead of four 64-bit ones per item
template <int MODE, typename IDX>
__global__ void __launch_bounds__(BW_THREADS, 3)
relu_maxpool_bwd_win_kernel(const uint4* __restrict__ ga, const uint4* __restrict__ c, int N, int H, int W, int C,
                            int Ho, int Wo, uint4* __restrict__ gc, float* __restrict__ dbias) {
  extern __shared__ float sacc[];  // [1][C]
  const int CV = C / 8;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cv = threadIdx.x % CV;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  using OFF = typename std::conditional<std::is_same<IDX, unsigned>::value, int, long long>::type;
  const int HP = (H + 1) / 2;                               // row pairs (the last one may be half)
  const int WP = MODE == 0 ? (W + 1) / 2 : W;               // column groups
  const IDX total = (IDX)N * HP * WP * CV;
  for (IDX chunk = blockIdx.x; chunk * (BW_THREADS * BW_ITER) < total; chunk += gridDim.x) {   // persistent blocks
  const IDX base = chunk * (BW_THREADS * BW_ITER);
  for (int it = 0; it < BW_ITER; ++it) {
    const IDX item = base + it * BW_THREADS + threadIdx.x;
    if (item >= total) break;
    IDX q = item / (IDX)CV;
    const IDX q2 = q / (IDX)WP;
    const int wg = (int)(q - q2 * (IDX)WP);
    const int n = (int)(q2 / (IDX)HP);
    const int hp = (int)(q2 - (IDX)n * (IDX)HP);
    const int h0 = 2 * hp, h1 = h0 + 1;
    const bool row1 = h1 < H;
    const bool win_h = hp < Ho;                              // a pooling window covers this row pair
    const uint4* cn = c + (long long)n * H * W * CV;
    uint4* gn = gc + (long long)n * H * W * CV;
    const uint4* gan = ga + (long long)n * Ho * Wo * CV;
    if (MODE == 0) {
      const int w0 = 2 * wg, w1 = w0 + 1;
      const bool col1 = w1 < W;
      const bool win = win_h && wg < Wo && row1 && col1;     // floor mode: partial windows do not exist
      uint4 v[4], g = make_uint4(0u, 0u, 0u, 0u), o[4];
      v[1] = v[2] = v[3] = g;
      const OFF i0 = ((OFF)h0 * W + w0) * CV + cv, rs = (OFF)W * CV;   // offsets inside one image
      v[0] = cn[i0];
      if (col1) v[1] = cn[i0 + CV];
      if (row1) v[2] = cn[i0 + rs];
      if (row1 && col1) v[3] = cn[i0 + rs + CV];
      if (win) g = gan[((OFF)hp * Wo + wg) * CV + cv];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bf2 v0 = as_bf2(u4c(v[0], k)), v1 = as_bf2(u4c(v[1], k)), v2 = as_bf2(u4c(v[2], k)), v3 = as_bf2(u4c(v[3], k));
        const bf2 m01 = __hmax2(v0, v1), m23 = __hmax2(v2, v3);
        // position k wins iff it is greater than every earlier and not smaller than every later one (first maximum)
        const bf2 p0 = __hge2(v0, __hmax2(v1, m23));
        const bf2 p1 = __hmul2(__hgt2(v1, v0), __hge2(v1, m23));
        const bf2 p2 = __hmul2(__hgt2(v2, m01), __hge2(v2, v3));
        const bf2 p3 = __hgt2(v3, __hmax2(m01, v2));
        const bf2 gm = __hmul2(as_bf2(u4c(g, k)), __hgt2(__hmax2(m01, m23), bf2_zero()));   // ReLU mask (c is post-ReLU)
        u4c(o[0], k) = as_u32(__hmul2(gm, p0)); u4c(o[1], k) = as_u32(__hmul2(gm, p1));
        u4c(o[2], k) = as_u32(__hmul2(gm, p2)); u4c(o[3], k) = as_u32(__hmul2(gm, p3));
        const float2 gf = __bfloat1622float2(gm);
        acc[0][2 * k] += gf.x; acc[0][2 * k + 1] += gf.y;
      }
      gn[i0] = o[0];
      if (col1) gn[i0 + CV] = o[1];
      if (row1) gn[i0 + rs] = o[2];
      if (row1 && col1) gn[i0 + rs + CV] = o[3];
    } else {
      const int w = wg;
      // columns w-1, w, w+1 of both rows; outside the image = padding = -inf (never a maximum)
      const uint4 NEG = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
      uint4 v[2][3], o[2], gl = make_uint4(0u, 0u, 0u, 0u), gr = gl;
      const OFF i0 = ((OFF)h0 * W + w) * CV + cv, rs = (OFF)W * CV;   // offsets inside one image
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int ww = w - 1 + d, hh = h0 + r;
          v[r][d] = (ww >= 0 && ww < W && hh < H) ? cn[i0 + r * rs + (d - 1) * CV] : NEG;
        }
      if (win_h && row1) {
        const OFF ig = ((OFF)hp * Wo + w) * CV + cv;
        gl = gan[ig];                                         // window wo = w   : columns w-1, w
        gr = gan[ig + CV];                                    // window wo = w+1 : columns w, w+1
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bf2 v00 = as_bf2(u4c(v[0][0], k)), v01 = as_bf2(u4c(v[0][1], k)), v02 = as_bf2(u4c(v[0][2], k));
        const bf2 v10 = as_bf2(u4c(v[1][0], k)), v11 = as_bf2(u4c(v[1][1], k)), v12 = as_bf2(u4c(v[1][2], k));
        // window wo = w, scan order (r0,w-1) (r0,w) (r1,w-1) (r1,w): positions 1 and 3 are ours
        const bf2 p1 = __hmul2(__hgt2(v01, v00), __hge2(v01, __hmax2(v10, v11)));
        const bf2 p3 = __hmul2(__hgt2(v11, __hmax2(v00, v01)), __hgt2(v11, v10));
        // window wo = w+1, scan order (r0,w) (r0,w+1) (r1,w) (r1,w+1): positions 0 and 2 are ours
        const bf2 q0 = __hge2(v01, __hmax2(v02, __hmax2(v11, v12)));
        const bf2 q2 = __hmul2(__hgt2(v11, __hmax2(v01, v02)), __hge2(v11, v12));
        const bf2 gL = as_bf2(u4c(gl, k)), gR = as_bf2(u4c(gr, k));
        bf2 o0 = __hfma2(gL, p1, __hmul2(gR, q0)), o1 = __hfma2(gL, p3, __hmul2(gR, q2));
        o0 = __hmul2(o0, __hgt2(v01, bf2_zero()));            // ReLU mask
        o1 = __hmul2(o1, __hgt2(v11, bf2_zero()));
        u4c(o[0], k) = as_u32(o0); u4c(o[1], k) = as_u32(o1);
        const float2 f0 = __bfloat1622float2(o0), f1 = __bfloat1622float2(o1);
        acc[0][2 * k] += f0.x + f1.x; acc[0][2 * k + 1] += f0.y + f1.y;
      }
      gn[i0] = o[0];
      if (row1) gn[i0 + rs] = o[1];
    }
  }
  }
  if (dbias) channel_reduce<1>(acc, cv, CV, C, sacc, dbias, 1);
}

// ---- stem backward ---------------------------------------------------------------------------------
// thread = (pooled pixel, 8-channel group); block-level reduction of dw[c][9], db[c] in shared memory
__global__ void __launch_bounds__(BW_THREADS)
hwr_stem_bwd_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b,
                    const uint4* __restrict__ ga, int N, int H, int W, int Cout, float* __restrict__ dw,
                    float* __restrict__ db) {
  extern __shared__ float sm[];  // w [Cout*9], b [Cout], acc [Cout*10]
  float* ws = sm;
  float* bs = sm + Cout * 9;
  float* accs = bs + Cout;
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = b[i];
  for (int i = threadIdx.x; i < Cout * 10; i += blockDim.x) accs[i] = 0.f;
  __syncthreads();
  const int CV = Cout / 8, Hp = H / 2, Wp = W / 2;
  const long long total = (long long)N * Hp * Wp * CV;
  const long long base = (long long)blockIdx.x * BW_THREADS * BW_ITER;
  const int cv = threadIdx.x % CV;
  float acc[8][10];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[j][k] = 0.f;
  for (int it = 0; it < BW_ITER; ++it) {
    const long long item = base + it * BW_THREADS + threadIdx.x;
    if (item >= total) break;
    long long pix = item / CV;
    const int wp = (int)(pix % Wp); pix /= Wp;
    const int hp = (int)(pix % Hp);
    const int n = (int)(pix / Hp);
    float patch[4][4];
    const float* im = img + (size_t)n * H * W;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int hh = 2 * hp - 1 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ww = 2 * wp - 1 + j;
        patch[i][j] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? im[(size_t)hh * W + ww] : 0.f;
      }
    }
    float gf[8];
    unpack8b(ga[item], gf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* k = ws + (cv * 8 + j) * 9;
      float m = -CUDART_INF_F; int by = 0, bx = 0;
#pragma unroll
      for (int oy = 0; oy < 2; ++oy)
#pragma unroll
        for (int ox = 0; ox < 2; ++ox) {
          float a = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) a = fmaf(k[ky * 3 + kx], patch[oy + ky][ox + kx], a);
          if (a > m) { m = a; by = oy; bx = ox; }   // first maximum wins, like max_pool2d
        }
      const float gv = (m + bs[cv * 8 + j] > 0.f) ? gf[j] : 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          // patch[by+ky][bx+kx] with by,bx in {0,1}: select without dynamic indexing
          const float p00 = patch[ky][kx], p01 = patch[ky][kx + 1], p10 = patch[ky + 1][kx], p11 = patch[ky + 1][kx + 1];
          const float pv = by ? (bx ? p11 : p10) : (bx ? p01 : p00);
          acc[j][ky * 3 + kx] = fmaf(gv, pv, acc[j][ky * 3 + kx]);
        }
      acc[j][9] += gv;
    }
  }
  // fold lanes that share cv, then shared atomics, then one global atomic per value per block
  if (CV < 32) {
    for (int off = 16; off >= CV; off >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[j][k] += __shfl_xor_sync(0xffffffffu, acc[j][k], off);
    }
  }
  const int lane = threadIdx.x & 31;
  if (CV >= 32 || lane < CV) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < 10; ++k) atomicAdd(&accs[(cv * 8 + j) * 10 + k], acc[j][k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 10; i += blockDim.x) {
    const int c = i / 10, k = i - c * 10;
    if (k < 9) atomicAdd(&dw[c * 9 + k], accs[i]);
    else atomicAdd(&db[c], accs[i]);
  }
}


// ---- generator: AdaIN + LeakyReLU (+ noise weight) backward ------------------------------------------
__global__ void __launch_bounds__(ST_THREADS)
adain_bwd_reduce_kernel(const uint4* __restrict__ g, const uint4* __restrict__ a, const float* __restrict__ save,
                        long long HW, int C, float* __restrict__ sums) {
  extern __shared__ unsigned char smraw[];
  const StreamSmem sm = stream_smem<2>(smraw);      // constants mean, rstd [2][C]
  const int n = blockIdx.y, CV = C / 8;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sm.cst[c] = save[((size_t)n * C + c) * 2]; sm.cst[C + c] = save[((size_t)n * C + c) * 2 + 1];
  }
  const int cv = threadIdx.x % CV;
  const float* kc = sm.cst + cv * 8;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const long long total = HW * CV;
  const uint4* const src[2] = {g + (size_t)n * total, a + (size_t)n * total};
  stream_chunks<2>(src, total, blockIdx.x, gridDim.x, sm.ring, sm.bars, [&](long long, const uint4 (&v)[2]) {
    float gf[8], af[8], mean[8], rstd[8];
    unpack8b(v[0], gf);
    unpack8b(v[1], af);
    ld8s(kc, mean); ld8s(kc + C, rstd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += gf[j];
      acc[1][j] += gf[j] * (af[j] - mean[j]) * rstd[j];
    }
  });
  channel_reduce_in<2>(acc, cv, CV, C, reinterpret_cast<float*>(sm.ring), sums + (size_t)n * C * 2, 2);
}

__global__ void __launch_bounds__(ST_THREADS)
adain_bwd_apply_kernel(const uint4* __restrict__ g, const uint4* __restrict__ a, const float* __restrict__ save,
                       const float* __restrict__ coef, const float* __restrict__ sums, int H, int W, int C,
                       float slope, const float* __restrict__ noise, unsigned long long seed,
                       unsigned long long subseq, const unsigned long long* __restrict__ seed_dev, int row_subseq,
                       uint4* __restrict__ gy, float* __restrict__ dch) {
  extern __shared__ unsigned char smraw[];
  const StreamSmem sm = stream_smem<2>(smraw);      // constants A, P, Q [3][C]
  const int n = blockIdx.y, CV = C / 8;
  const long long HW = (long long)H * W;
  const float inv = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // ga = A*(g - k0 - ahat*k1), ahat = (a - mean)*rstd   ==   A*g + P*a + Q
    const size_t i = ((size_t)n * C + c) * 2;
    const float mean = save[i], rstd = save[i + 1], A = coef[i];
    const float k0 = sums[i] * inv, k1 = sums[i + 1] * inv;
    sm.cst[c] = A;
    sm.cst[C + c] = -A * k1 * rstd;
    sm.cst[2 * C + c] = A * (k1 * rstd * mean - k0);
  }
  const int cv = threadIdx.x % CV;
  const float* kc = sm.cst + cv * 8;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  seed += seed_dev ? *seed_dev : 0ull;
  const uint2 nkey = noise_key(seed, subseq);
  const long long total = HW * CV;
  uint4* yn = gy + (size_t)n * total;
  const uint4* const src[2] = {g + (size_t)n * total, a + (size_t)n * total};
  stream_chunks<2>(src, total, blockIdx.x, gridDim.x, sm.ring, sm.bars, [&](long long item, const uint4 (&v)[2]) {
    float gf[8], af[8], o[8], z[8], A[8], P[8], Q[8];
    unpack8b(v[0], gf);
    unpack8b(v[1], af);
    if (noise) {
      const float4* zp = reinterpret_cast<const float4*>(noise + ((size_t)n * total + item) * 8);
      const float4 z0 = zp[0], z1 = zp[1];
      z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
    } else {
      unsigned long long e;   // element index of channel cv*8 in the forward launch's numbering
      uint2 key = nkey;
      if (row_subseq) {
        const long long pix = item / CV;
        const int h = (int)(pix / W), w = (int)(pix - (long long)h * W);
        key = noise_key(seed, subseq + (unsigned long long)h);
        e = ((unsigned long long)n * W + w) * (unsigned long long)C + cv * 8;
      } else {
        e = ((unsigned long long)n * total + item) * 8ull;
      }
      normal_oct(key, e >> 1, z);   // e is a multiple of 8
    }
    ld8s(kc, A); ld8s(kc + C, P); ld8s(kc + 2 * C, Q);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ga = fmaf(A[j], gf[j], fmaf(P[j], af[j], Q[j]));
      o[j] = af[j] > 0.f ? ga : ga * slope;
      acc[0][j] += o[j];
      acc[1][j] = fmaf(o[j], z[j], acc[1][j]);
    }
    yn[item] = pack8b(o);
  });
  channel_reduce_in<2>(acc, cv, CV, C, reinterpret_cast<float*>(sm.ring), dch, 2);
}

// ---- generator output backward ----------------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS)
gen_output_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ out, const uint4* __restrict__ a,
                      const float* __restrict__ coef, const float* __restrict__ w, long long HW, int C,
                      uint4* __restrict__ gx, float* __restrict__ dwb) {
  extern __shared__ float sm[];  // A[C], B[C], w[C], acc[C+1]
  const int n = blockIdx.y, CV = C / 8;
  float* As = sm; float* Bs = sm + C; float* ws = sm + 2 * C; float* accs = sm + 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    As[c] = coef[((size_t)n * C + c) * 2]; Bs[c] = coef[((size_t)n * C + c) * 2 + 1]; ws[c] = w[c];
  }
  for (int c = threadIdx.x; c <= C; c += blockDim.x) accs[c] = 0.f;
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = pix < HW;
  float gp = 0.f;
  if (valid) {
    const float o = out[(size_t)n * HW + pix];
    gp = g_out[(size_t)n * HW + pix] * (1.f - o * o);
  }
  const uint4* ap = a + ((size_t)n * HW + (valid ? pix : 0)) * CV;
  uint4* gp_out = gx + ((size_t)n * HW + (valid ? pix : 0)) * CV;
  for (int v = 0; v < CV; ++v) {   // every lane runs the same warp reductions; tail lanes contribute zeros
    float af[8], o8[8];
    unpack8b(ap[v], af);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      o8[j] = ws[c] * gp;
      // dw[c] += gp * x_last[c]: warp-reduce, then one shared atomic per warp
      const float t = warp_sum(gp * fmaf(As[c], af[j], Bs[c]));
      if ((threadIdx.x & 31) == 0) atomicAdd(&accs[c], t);
    }
    if (valid) gp_out[v] = pack8b(o8);
  }
  const float s = warp_sum(gp);
  if ((threadIdx.x & 31) == 0) atomicAdd(&accs[C], s);
  __syncthreads();
  for (int c = threadIdx.x; c <= C; c += blockDim.x) atomicAdd(&dwb[c], accs[c]);
}

// The same pass for C = 8*CV <= 32 channels (the generator's last block has 16) with the weight-gradient sums kept in
// registers: a thread walks a strided set of pixels and adds gp * x_last[c] into its own C accumulators; the warps are
// folded once per block.  The generic kernel above folds every channel of every pixel over the warp as it goes
// (5 shuffles + a shared atomic per channel and pixel): 283 us for 555 MB at 128 lines (ncu, round 2: short_scoreboard 59 %).
template <int CV>
__global__ void __launch_bounds__(BW_THREADS)
gen_output_bwd_acc_kernel(const float* __restrict__ g_out, const float* __restrict__ out, const uint4* __restrict__ a,
                          const float* __restrict__ coef, const float* __restrict__ w, long long HW,
                          uint4* __restrict__ gx, float* __restrict__ dwb) {
  constexpr int C = CV * 8;
  __shared__ float As[C], Bs[C], ws[C], red[C + 1];
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    As[c] = coef[((size_t)n * C + c) * 2]; Bs[c] = coef[((size_t)n * C + c) * 2 + 1]; ws[c] = w[c];
  }
  for (int c = threadIdx.x; c <= C; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  float dacc[C], sg = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) dacc[c] = 0.f;
  const float* gon = g_out + (size_t)n * HW;
  const float* on = out + (size_t)n * HW;
  const uint4* an = a + (size_t)n * HW * CV;
  uint4* gn = gx + (size_t)n * HW * CV;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += (long long)gridDim.x * blockDim.x) {
    const float o = on[pix];
    const float gp = gon[pix] * (1.f - o * o);
    sg += gp;
    uint4 av[CV];
#pragma unroll
    for (int v = 0; v < CV; ++v) av[v] = ldg_stream(an + pix * CV + v);
#pragma unroll
    for (int v = 0; v < CV; ++v) {
      float af[8], o8[8];
      unpack8b(av[v], af);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = v * 8 + j;
        o8[j] = ws[c] * gp;
        dacc[c] = fmaf(gp, fmaf(As[c], af[j], Bs[c]), dacc[c]);
      }
      gn[pix * CV + v] = pack8b(o8);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float t = warp_sum(dacc[c]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[c], t);
  }
  sg = warp_sum(sg);
  if ((threadIdx.x & 31) == 0) atomicAdd(&red[C], sg);
  __syncthreads();
  for (int c = threadIdx.x; c <= C; c += blockDim.x) atomicAdd(&dwb[c], red[c]);
}

// thread = (pooled pixel, 8-channel group): writes the four conv0-output gradients of its 2x2 window
__global__ void __launch_bounds__(BW_THREADS)
hwr_stem_bwd_expand_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b,
                           const uint4* __restrict__ ga, int N, int H, int W, int Cout, uint4* __restrict__ gc0) {
  extern __shared__ float sm[];  // w [Cout*9], b [Cout]
  float* ws = sm;
  float* bs = sm + Cout * 9;
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = b[i];
  __syncthreads();
  const int CV = Cout / 8, Hp = H / 2, Wp = W / 2;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)N * Hp * Wp * CV) return;
  const int cv = (int)(item % CV);
  long long pix = item / CV;
  const int wp = (int)(pix % Wp); pix /= Wp;
  const int hp = (int)(pix % Hp);
  const int n = (int)(pix / Hp);
  float patch[4][4];
  const float* im = img + (size_t)n * H * W;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int hh = 2 * hp - 1 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ww = 2 * wp - 1 + j;
      patch[i][j] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? im[(size_t)hh * W + ww] : 0.f;
    }
  }
  float gf[8], o[4][8];
  unpack8b(ga[item], gf);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* k = ws + (cv * 8 + j) * 9;
    float m = -CUDART_INF_F; int best = 0;
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
      for (int ox = 0; ox < 2; ++ox) {
        float a = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) a = fmaf(k[ky * 3 + kx], patch[oy + ky][ox + kx], a);
        if (a > m) { m = a; best = oy * 2 + ox; }
      }
    const float gv = (m + bs[cv * 8 + j] > 0.f) ? gf[j] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q][j] = (q == best) ? gv : 0.f;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int hh = 2 * hp + (q >> 1), ww = 2 * wp + (q & 1);
    gc0[(((size_t)n * H + hh) * W + ww) * CV + cv] = pack8b(o[q]);
  }
}

// ---- stem: image gradient in one pass ------------------------------------------------------------------
// d loss / d image for conv0 (1->Cout, 3x3, pad 1) + ReLU + MaxPool 2x2 (cnn_only_hwr.py:44-46), needed by the GAN
// lessons (the recognizer reads generated lines).  Replaces hwr_stem_bwd_expand (which materialises the
// [N,H,W,Cout] bf16 gradient of conv0's output: 134 MB at 16 lines) + a 9-tap tensor-core dgrad with one output
// channel + a channel-0 slice copy.  A thread owns (pooled pixel, 8 channels): it recomputes the four conv0 outputs of
// its window, finds the first maximum (ATen's rule), and scatters gy * w[c][ky][kx] of the winner into a private
// 4x4 image patch (the window plus its one-pixel ring); the eight channel groups of a pooled pixel are folded with a
// 14-shuffle recursive halving, accumulated into a shared-memory image tile (8 pooled rows x 32 pooled columns per
// block) and flushed with one fp32 global atomic per image pixel of the tile.
constexpr int SB_PR = 8, SB_PC = 32;          // pooled rows / columns per block
__global__ void __launch_bounds__(256)
hwr_stem_bwd_image_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b,
                          const uint4* __restrict__ ga, int N, int H, int W, float* __restrict__ gimg) {
  constexpr int COUT = 64, CV = 8;
  constexpr int TR = 2 * SB_PR + 2, TC = 2 * SB_PC + 2;
  __shared__ float ws[COUT * 9], bs[COUT];
  __shared__ float tile[TR * TC];
  for (int i = threadIdx.x; i < COUT * 9; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = b[i];
  for (int i = threadIdx.x; i < TR * TC; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  const int Hp = H / 2, Wp = W / 2;
  const int n = blockIdx.z, hp0 = blockIdx.y * SB_PR, wp0 = blockIdx.x * SB_PC;
  const int cv = threadIdx.x & 7, px = threadIdx.x >> 3;
  const int wp = wp0 + px;
  const float* im = img + (size_t)n * H * W;
  for (int r = 0; r < SB_PR; ++r) {
    const int hp = hp0 + r;
    if (hp >= Hp) break;                                  // uniform over the block
    float gp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) gp[i] = 0.f;
    if (wp < Wp) {
      float patch[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int hh = 2 * hp - 1 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ww = 2 * wp - 1 + j;
          patch[i][j] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? im[(size_t)hh * W + ww] : 0.f;
        }
      }
      float gf[8];
      unpack8b(ga[(((size_t)n * Hp + hp) * Wp + wp) * CV + cv], gf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* k = ws + (cv * 8 + j) * 9;
        float kk[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) kk[t] = k[t];
        float m = -CUDART_INF_F; int best = 0;
#pragma unroll
        for (int oy = 0; oy < 2; ++oy)
#pragma unroll
          for (int ox = 0; ox < 2; ++ox) {
            float a = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) a = fmaf(kk[ky * 3 + kx], patch[oy + ky][ox + kx], a);
            if (a > m) { m = a; best = oy * 2 + ox; }
          }
        const float gv = (m + bs[cv * 8 + j] > 0.f) ? gf[j] : 0.f;
        // conv0 output pixel (oy, ox) of the window reads patch[oy+ky][ox+kx]: its gradient lands there
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float gq = (q == best) ? gv : 0.f;
          const int oy = q >> 1, ox = q & 1;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) gp[(oy + ky) * 4 + ox + kx] = fmaf(gq, kk[ky * 3 + kx], gp[(oy + ky) * 4 + ox + kx]);
        }
      }
    }
    // fold the 8 channel groups (lanes cv = 0..7 of the same pooled pixel): afterwards lane cv holds entries 2cv, 2cv+1
#pragma unroll
    for (int half = 8; half >= 2; half >>= 1) {
      const bool upper = (cv & (half >> 1)) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = upper ? gp[i] : gp[i + half];
        const float keep = upper ? gp[i + half] : gp[i];
        gp[i] = keep + __shfl_xor_sync(0xffffffffu, send, half >> 1);
      }
    }
    if (wp < Wp) {
      const int e0 = 2 * cv, pi = e0 >> 2, pj = e0 & 3;   // patch row / first column of this lane's two entries
      float* t = &tile[(2 * r + pi) * TC + 2 * px + pj];
      atomicAdd(t, gp[0]);
      atomicAdd(t + 1, gp[1]);
    }
  }
  __syncthreads();
  float* gn = gimg + (size_t)n * H * W;
  for (int i = threadIdx.x; i < TR * TC; i += blockDim.x) {
    const int tr = i / TC, tc = i - tr * TC;
    const int hh = 2 * hp0 - 1 + tr, ww = 2 * wp0 - 1 + tc;
    const float v = tile[i];
    if (hh >= 0 && hh < H && ww >= 0 && ww < W && v != 0.f) atomicAdd(&gn[(size_t)hh * W + ww], v);
  }
}

static inline unsigned bw_blocks(long long items, int per_block) {
  return (unsigned)((items + per_block - 1) / per_block);
}
// persistent streaming kernels: at most `per_sm` blocks per SM (148 SMs), each looping over chunks, so that the
// per-block prologue (constants -> shared memory) and epilogue (shared -> global atomics) are paid ~600 times per
// launch instead of once per 64 KB of input
static inline unsigned bw_blocks_persistent(long long items, int per_block, int split = 1, int per_sm = 2) {
  const long long need = (items + per_block - 1) / per_block;
  const long long cap = (148LL * per_sm + split - 1) / split;
  return (unsigned)(need < cap ? need : cap);
}
static inline bool cv_ok(int C) {
  const int cv = C / 8;
  return C % 8 == 0 && cv > 0 && (cv & (cv - 1)) == 0 && cv <= BW_THREADS;
}

}  // namespace hwg

using namespace hwg;

extern "C" int hwg_logsoftmax_bwd(const float* g, const float* lp, int T, int B, int C, int Cp, void* gz,
                                  float* dbias, void* stream) {
  HWG_REQUIRE(g && lp && gz && dbias && T > 0 && B > 0 && C > 0 && Cp >= C, "hwg_logsoftmax_bwd: bad argument");
  const long long rows = (long long)T * B;
  logsoftmax_bwd_kernel<<<bw_blocks(rows, 8), 256, (size_t)C * sizeof(float), (cudaStream_t)stream>>>(
      g, lp, T, B, C, Cp, reinterpret_cast<__nv_bfloat16*>(gz), dbias);
  return check_launch("logsoftmax_bwd_kernel");
}

extern "C" int hwg_bn_bwd_reduce(const void* g, const void* z, const float* coef, const float* save, int64_t rows,
                                 int C, int relu, float* sums, void* stream) {
  HWG_REQUIRE(g && z && coef && save && sums && rows > 0, "hwg_bn_bwd_reduce: bad argument");
  HWG_REQUIRE(cv_ok(C), "hwg_bn_bwd_reduce: C=%d must be 8 x a power of two", C);
  const long long total = rows * (C / 8);
  HWG_SMEM_OPTIN(bn_bwd_reduce_kernel);
  bn_bwd_reduce_kernel<<<bw_blocks_persistent(total, ST_CHUNK, 1, 3), ST_THREADS, stream_kernel_smem(2, 4 * C),
                         (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(z),
                                                 coef, save, rows, C, relu, sums);
  return check_launch("bn_bwd_reduce_kernel");
}

extern "C" int hwg_bn_bwd_apply(const void* g, const void* z, const float* coef, const float* save,
                                const float* weight, const float* sums, int64_t rows, int64_t norm_rows, int C,
                                int relu, void* gz, float* dconv_bias, void* stream) {
  HWG_REQUIRE(g && z && coef && save && sums && gz && rows > 0, "hwg_bn_bwd_apply: bad argument");
  HWG_REQUIRE(cv_ok(C), "hwg_bn_bwd_apply: C=%d must be 8 x a power of two", C);
  const long long total = rows * (C / 8);
  HWG_SMEM_OPTIN(bn_bwd_apply_kernel);
  bn_bwd_apply_kernel<<<bw_blocks_persistent(total, ST_CHUNK, 1, 3), ST_THREADS, stream_kernel_smem(2, 5 * C),
                        (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(z),
                                                coef, save, weight, sums, rows, norm_rows > 0 ? norm_rows : rows, C,
                                                relu, reinterpret_cast<uint4*>(gz), dconv_bias);
  return check_launch("bn_bwd_apply_kernel");
}

extern "C" int hwg_relu_maxpool_bwd(const void* ga, const void* c, int N, int H, int W, int C, int kh, int kw,
                                    int sh, int sw, int ph, int pw, int Ho, int Wo, void* gc, float* dbias,
                                    void* stream) {
  HWG_REQUIRE(ga && c && gc && N > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0, "hwg_relu_maxpool_bwd: bad argument");
  HWG_REQUIRE(cv_ok(C), "hwg_relu_maxpool_bwd: C=%d must be 8 x a power of two", C);
  HWG_REQUIRE(Ho == (H + 2 * ph - kh) / sh + 1 && Wo == (W + 2 * pw - kw) / sw + 1,
              "hwg_relu_maxpool_bwd: Ho/Wo do not match the pooling geometry");
  if (kh == 2 && kw == 2 && sh == 2 && ph == 0 && ((sw == 2 && pw == 0) || (sw == 1 && pw == 1)) && cv_ok(C)) {
    const int mode = sw == 2 ? 0 : 1;
    const long long items = (long long)N * ((H + 1) / 2) * (mode == 0 ? (W + 1) / 2 : W) * (C / 8);
    // 32-bit item indices and per-image offsets (IDX = unsigned, OFF = int) when both fit
    const bool small = items < (1LL << 31) - (long long)BW_THREADS * BW_ITER * 148 * 4 &&
                       (long long)H * W * (C / 8) < (1LL << 31);
    auto kern = small ? (mode == 0 ? relu_maxpool_bwd_win_kernel<0, unsigned> : relu_maxpool_bwd_win_kernel<1, unsigned>)
                      : (mode == 0 ? relu_maxpool_bwd_win_kernel<0, long long> : relu_maxpool_bwd_win_kernel<1, long long>);
    kern<<<bw_blocks_persistent(items, BW_THREADS * BW_ITER, 1, 3), BW_THREADS, (size_t)C * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(ga), reinterpret_cast<const uint4*>(c), N, H, W, C, Ho, Wo,
        reinterpret_cast<uint4*>(gc), dbias);
    return check_launch("relu_maxpool_bwd_win_kernel");
  }
  const long long total = (long long)N * H * W * (C / 8);
  relu_maxpool_bwd_kernel<<<bw_blocks(total, BW_THREADS * BW_ITER), BW_THREADS, (size_t)C * sizeof(float),
                            (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(ga),
                                                    reinterpret_cast<const uint4*>(c), N, H, W, C, kh, kw, sh, sw, ph,
                                                    pw, Ho, Wo, reinterpret_cast<uint4*>(gc), dbias);
  return check_launch("relu_maxpool_bwd_kernel");
}

extern "C" int hwg_hwr_stem_bwd(const float* img, const float* w, const float* b, const void* ga, int N, int H,
                                int W, int Cout, float* dw, float* db, void* stream) {
  HWG_REQUIRE(img && w && b && ga && dw && db && N > 0, "hwg_hwr_stem_bwd: bad argument");
  HWG_REQUIRE(H % 2 == 0 && W % 2 == 0 && cv_ok(Cout), "hwg_hwr_stem_bwd: bad geometry");
  const long long total = (long long)N * (H / 2) * (W / 2) * (Cout / 8);
  hwr_stem_bwd_kernel<<<bw_blocks(total, BW_THREADS * BW_ITER), BW_THREADS, (size_t)Cout * 20 * sizeof(float),
                        (cudaStream_t)stream>>>(img, w, b, reinterpret_cast<const uint4*>(ga), N, H, W, Cout, dw, db);
  return check_launch("hwr_stem_bwd_kernel");
}

extern "C" int hwg_adain_bwd_reduce(const void* g, const void* a, const float* save, int N, int64_t HW, int C,
                                    float* sums, void* stream) {
  HWG_REQUIRE(g && a && save && sums && N > 0 && HW > 0, "hwg_adain_bwd_reduce: bad argument");
  HWG_REQUIRE(cv_ok(C), "hwg_adain_bwd_reduce: C=%d must be 8 x a power of two", C);
  dim3 grid(bw_blocks_persistent(HW * (C / 8), ST_CHUNK, N, 3), N);
  HWG_SMEM_OPTIN(adain_bwd_reduce_kernel);
  adain_bwd_reduce_kernel<<<grid, ST_THREADS, stream_kernel_smem(2, 2 * C), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(a), save, HW, C, sums);
  return check_launch("adain_bwd_reduce_kernel");
}

extern "C" int hwg_adain_bwd_apply(const void* g, const void* a, const float* save, const float* coef,
                                   const float* sums, int N, int H, int W, int C, float slope, const float* noise,
                                   uint64_t noise_seed, uint64_t noise_subseq, const uint64_t* noise_seed_dev,
                                   int row_subseq, void* gy, float* dch, void* stream) {
  HWG_REQUIRE(g && a && save && coef && sums && gy && dch && N > 0 && H > 0 && W > 0, "hwg_adain_bwd_apply: bad argument");
  HWG_REQUIRE(cv_ok(C), "hwg_adain_bwd_apply: C=%d must be 8 x a power of two", C);
  dim3 grid(bw_blocks_persistent((long long)H * W * (C / 8), ST_CHUNK, N, 3), N);
  HWG_SMEM_OPTIN(adain_bwd_apply_kernel);
  adain_bwd_apply_kernel<<<grid, ST_THREADS, stream_kernel_smem(2, 3 * C), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(a), save, coef, sums, H, W, C, slope, noise,
      noise_seed, noise_subseq, reinterpret_cast<const unsigned long long*>(noise_seed_dev), row_subseq,
      reinterpret_cast<uint4*>(gy), dch);
  return check_launch("adain_bwd_apply_kernel");
}

extern "C" int hwg_gen_output_bwd(const float* g_out, const float* out, const void* a, const float* coef,
                                  const float* w, int N, int64_t HW, int C, void* gx, float* dwb, void* stream) {
  HWG_REQUIRE(g_out && out && a && coef && w && gx && dwb && N > 0 && HW > 0 && C % 8 == 0,
              "hwg_gen_output_bwd: bad argument");
  if (C == 16 || C == 32) {
    // a few blocks per image, each thread keeping its sums over ~HW / (blocks x 256) pixels
    long long bx = (148LL * 8 + N - 1) / N;
    const long long need = bw_blocks(HW, BW_THREADS);
    if (bx > need) bx = need;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, N);
    if (C == 16)
      gen_output_bwd_acc_kernel<2><<<grid, BW_THREADS, 0, (cudaStream_t)stream>>>(
          g_out, out, reinterpret_cast<const uint4*>(a), coef, w, HW, reinterpret_cast<uint4*>(gx), dwb);
    else
      gen_output_bwd_acc_kernel<4><<<grid, BW_THREADS, 0, (cudaStream_t)stream>>>(
          g_out, out, reinterpret_cast<const uint4*>(a), coef, w, HW, reinterpret_cast<uint4*>(gx), dwb);
    return check_launch("gen_output_bwd_acc_kernel");
  }
  dim3 grid(bw_blocks(HW, BW_THREADS), N);
  gen_output_bwd_kernel<<<grid, BW_THREADS, (size_t)(4 * C + 1) * sizeof(float), (cudaStream_t)stream>>>(
      g_out, out, reinterpret_cast<const uint4*>(a), coef, w, HW, C, reinterpret_cast<uint4*>(gx), dwb);
  return check_launch("gen_output_bwd_kernel");
}

extern "C" int hwg_hwr_stem_bwd_image(const float* img, const float* w, const float* b, const void* ga, int N, int H,
                                      int W, int Cout, float* gimg, void* stream) {
  HWG_REQUIRE(img && w && b && ga && gimg && N > 0, "hwg_hwr_stem_bwd_image: bad argument");
  HWG_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cout == 64, "hwg_hwr_stem_bwd_image: needs even H, W and Cout == 64");
  HWG_REQUIRE(N <= 65535, "hwg_hwr_stem_bwd_image: N too large");
  dim3 grid((W / 2 + SB_PC - 1) / SB_PC, (H / 2 + SB_PR - 1) / SB_PR, N);
  hwr_stem_bwd_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, w, b, reinterpret_cast<const uint4*>(ga), N, H,
                                                                   W, gimg);
  return check_launch("hwr_stem_bwd_image_kernel");
}

extern "C" int hwg_hwr_stem_bwd_expand(const float* img, const float* w, const float* b, const void* ga, int N,
                                       int H, int W, int Cout, void* gc0, void* stream) {
  HWG_REQUIRE(img && w && b && ga && gc0 && N > 0, "hwg_hwr_stem_bwd_expand: bad argument");
  HWG_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cout % 8 == 0, "hwg_hwr_stem_bwd_expand: bad geometry");
  const long long total = (long long)N * (H / 2) * (W / 2) * (Cout / 8);
  hwr_stem_bwd_expand_kernel<<<bw_blocks(total, BW_THREADS), BW_THREADS, (size_t)Cout * 10 * sizeof(float),
                               (cudaStream_t)stream>>>(img, w, b, reinterpret_cast<const uint4*>(ga), N, H, W, Cout,
                                                       reinterpret_cast<uint4*>(gc0));
  return check_launch("hwr_stem_bwd_expand_kernel");
}
