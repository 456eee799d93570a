// HBM-bound fused passes around the convolutions (sm_100a): every kernel moves 16-byte
// vectors of 8 bf16 channels, NHWC, and touches each activation once.
#include "common.cuh"
#include "stream.cuh"
#include "noise_rng.cuh"
#include <math_constants.h>

namespace hwg {
// flat item index -> (a, b, c, d) with item = ((d * C + c) * B + b) * A + a; 32-bit divisions when the index fits
__device__ __forceinline__ void decode4(long long item, int A, int B, int C, int& a, int& b, int& c, int& d) {
  if (item < (1LL << 31)) {
    const unsigned i = (unsigned)item, q1 = i / (unsigned)A, q2 = q1 / (unsigned)B, q3 = q2 / (unsigned)C;
    a = (int)(i - q1 * (unsigned)A); b = (int)(q1 - q2 * (unsigned)B); c = (int)(q2 - q3 * (unsigned)C); d = (int)q3;
  } else {
    const long long q1 = item / A, q2 = q1 / B, q3 = q2 / C;
    a = (int)(item - q1 * A); b = (int)(q1 - q2 * B); c = (int)(q2 - q3 * C); d = (int)q3;
  }
}


__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ float act_fn(float v, int act, float slope) {
  if (act == HWG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == HWG_ACT_LRELU) return v > 0.f ? v : v * slope;
  return v;
}

constexpr int EW_THREADS = 256;
constexpr int EW_ITER = 8;  // vectors per thread per block

// ---- small dense layers ---------------------------------------------------------------------
__global__ void linear_f32_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                  const float* __restrict__ bias, float* __restrict__ out, int B, int K,
                                  int O, int act, float slope) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * O) return;
  const int b = gw / O, o = gw - b * O;
  const float* x = in + (size_t)b * K;
  const float* w = W + (size_t)o * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(x[k], w[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)b * O + o] = act_fn(acc + (bias ? bias[o] : 0.f), act, slope);
}

__global__ void pixelnorm_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int K) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B) return;
  const float* x = in + (size_t)gw * K;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += x[k] * x[k];
  s = warp_sum(s);
  const float r = rsqrtf(s / (float)K + 1e-8f);
  for (int k = lane; k < K; k += 32) out[(size_t)gw * K + k] = x[k] * r;
}

__global__ void gen_pack_input_kernel(const float* __restrict__ content, long long cs_t, long long cs_b,
                                      long long cs_c, const float* __restrict__ style, int T, int B, int C,
                                      int S, int Cp, uint4* __restrict__ x) {
  const int CV = Cp / 8;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)B * T * CV) return;
  const int cv = (int)(item % CV);
  const long long pix = item / CV;
  const int t = (int)(pix % T), b = (int)(pix / T);
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    float v = 0.f;
    if (c < C) v = content[t * cs_t + b * cs_b + c * cs_c];
    else if (c < C + S) v = style[(size_t)b * S + (c - C)];
    f[j] = v;
  }
  x[item] = pack8(f);
}

// ---- normalisation coefficients -------------------------------------------------------------
__global__ void adain_coeffs_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, long long gbs, int N, int C, float inv_hw,
                                    float eps, float* __restrict__ coef, float* __restrict__ save) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  const float mean = stats[2 * i] * inv_hw;
  const float var = fmaxf(stats[2 * i + 1] * inv_hw - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float a = gamma[n * gbs + c] * rstd;
  coef[2 * i] = a;
  coef[2 * i + 1] = beta[n * gbs + c] - mean * a;
  if (save) { save[2 * i] = mean; save[2 * i + 1] = rstd; }
}

__global__ void bn_coeffs_kernel(const float* __restrict__ stats, int N, int C, float count,
                                 const float* __restrict__ weight, const float* __restrict__ bias,
                                 float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                 int use_batch, float* __restrict__ coef, float* __restrict__ save) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (use_batch) {
    float s1 = 0.f, s2 = 0.f;
    for (int n = 0; n < N; ++n) { s1 += stats[((size_t)n * C + c) * 2]; s2 += stats[((size_t)n * C + c) * 2 + 1]; }
    mean = s1 / count;
    var = fmaxf(s2 / count - mean * mean, 0.f);
    if (rmean) {
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * var * (count / fmaxf(count - 1.f, 1.f));
    }
  } else {
    mean = rmean[c]; var = rvar[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float a = (weight ? weight[c] : 1.f) * rstd;
  coef[2 * c] = a;
  coef[2 * c + 1] = (bias ? bias[c] : 0.f) - mean * a;
  if (save) { save[2 * c] = mean; save[2 * c + 1] = rstd; }
}

// ---- y = act(a*x+b) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
scale_shift_act_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ coef,
                       int per_sample, long long HW, int C, int act, float slope) {
  extern __shared__ float cs[];  // [C][2]
  const int n = blockIdx.y, CV = C / 8;
  const float* cf = coef + (per_sample ? (size_t)n * C * 2 : 0);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) cs[i] = cf[i];
  __syncthreads();
  const long long total = HW * CV;
  const uint4* xn = x + (size_t)n * total;
  uint4* yn = y + (size_t)n * total;
  constexpr int U = 4;   // independent 16-byte loads issued before any is consumed (bytes in flight per thread)
  for (long long chunk = blockIdx.x; chunk * (EW_THREADS * EW_ITER) < total; chunk += gridDim.x)   // persistent blocks
  for (int it0 = 0; it0 < EW_ITER; it0 += U) {
    const long long base = chunk * (EW_THREADS * EW_ITER) + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long item = base + (long long)(it0 + u) * EW_THREADS;
      if (item < total) v[u] = xn[item];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long item = base + (long long)(it0 + u) * EW_THREADS;
      if (item >= total) continue;
      const int cv = (int)(item % CV);
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 ab = *reinterpret_cast<const float2*>(&cs[2 * (cv * 8 + j)]);
        f[j] = act_fn(fmaf(ab.x, f[j], ab.y), act, slope);
      }
      yn[item] = pack8(f);
    }
  }
}

// Bulk-staged variant (stream.cuh) for channel counts 8 x 2^k: the TMA engine streams x through a shared-memory ring,
// the per-channel (a, b) pairs sit in shared memory as two planes (conflict-free: lanes read consecutive words).
__global__ void __launch_bounds__(ST_THREADS)
scale_shift_act_stream_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ coef,
                              int per_sample, long long HW, int C, int act, float slope) {
  extern __shared__ unsigned char smraw[];
  unsigned char* ring = smraw + ((128u - (sm100::smem_u32(smraw) & 127u)) & 127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)ST_STAGES * ST_CHUNK * 16);
  float* cst = reinterpret_cast<float*>(bars + ST_BAR_SLOTS);     // a [C], b [C]; 16-byte aligned
  const int n = blockIdx.y, CV = C / 8;
  const float* cf = coef + (per_sample ? (size_t)n * C * 2 : 0);
  for (int c = threadIdx.x; c < C; c += blockDim.x) { cst[c] = cf[2 * c]; cst[C + c] = cf[2 * c + 1]; }
  const float* kc = cst + (threadIdx.x % CV) * 8;
  const long long total = HW * CV;
  uint4* yn = y + (size_t)n * total;
  const uint4* const src[1] = {x + (size_t)n * total};
  stream_chunks<1>(src, total, blockIdx.x, gridDim.x, ring, bars, [&](long long item, const uint4 (&v)[1]) {
    float f[8], a[8], b[8];
    unpack8(v[0], f);
    ld8s(kc, a); ld8s(kc + C, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_fn(fmaf(a[j], f[j], b[j]), act, slope);
    yn[item] = pack8(f);
  });
}

// ---- blur + noise + activation + statistics ---------------------------------------------------
// A block owns BLUR_ROWS rows x EW_THREADS 16-byte items of one line.  One thread pulls the tile plus its halo
// (one row above / below, one pixel left / right) into shared memory with 1-D bulk copies, one per row
// (cp.async.bulk, completion on an mbarrier): the whole tile is in flight at once — a register-only rolling window
// kept three loads per thread outstanding and measured 2.2 TB/s.  A thread then owns one channel vector of one
// image column and walks the rows downwards with a separable rolling window (l + 2c + r per row, combined over the
// rows above / at / below), all from shared memory.
// ncu on that version (round 2): 76-83 % issue activity at 275 (plain blur) / 437 (with noise) instructions per item —
// the edge selects of every tap, the runtime activation / statistics / noise options and the register rotation of the
// rolled row loop.  Now the halo OUTSIDE the image is zero-filled in shared memory by the border blocks (no selects in
// the taps), the row loop is unrolled (the rotation is renaming) and the two shapes of the hot path are compile-time
// MODEs: 1 = generator forward (in-kernel noise, LeakyReLU, statistics), 2 = adjoint in the backward (plain blur);
// 0 keeps every option at run time.
constexpr int BLUR_ROWS = 8;

__device__ __forceinline__ void blur_hrow(const uint4* __restrict__ p, int CV, float (&hb)[8]) {
  float c[8], l[8], q[8];
  unpack8(p[0], c);
  unpack8(p[-CV], l);
  unpack8(p[CV], q);
#pragma unroll
  for (int j = 0; j < 8; ++j) hb[j] = fmaf(2.f, c[j], l[j] + q[j]);
}

__host__ __device__ constexpr size_t blur_tile_bytes(int CV) { return (size_t)(BLUR_ROWS + 2) * (EW_THREADS + 2 * CV) * 16; }

template <int MODE>
__global__ void __launch_bounds__(EW_THREADS, 3)
blur_noise_act_stats_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int C, int cv_shift,
                            const float* __restrict__ noise, const float* __restrict__ noise_w,
                            unsigned long long seed, unsigned long long subseq,
                            const unsigned long long* __restrict__ seed_dev, int act, float slope,
                            float* __restrict__ stats) {
  extern __shared__ __align__(128) unsigned char blur_smem[];
  const bool has_noise = MODE == 1 || (MODE == 0 && noise_w != nullptr);
  const bool use_rng = MODE == 1 || (MODE == 0 && noise == nullptr);
  const bool has_stats = MODE == 1 || (MODE == 0 && stats != nullptr);
  const int actv = MODE == 1 ? HWG_ACT_LRELU : (MODE == 2 ? HWG_ACT_NONE : act);
  const int n = blockIdx.z, CV = C / 8, pitch = EW_THREADS + 2 * CV;
  uint4* tile = reinterpret_cast<uint4*>(blur_smem);                      // [BLUR_ROWS + 2][pitch]
  float* sacc = reinterpret_cast<float*>(blur_smem + blur_tile_bytes(CV));   // [C][2] block-level statistics
  uint64_t* bar = reinterpret_cast<uint64_t*>(sacc + 2 * C);
  const long long row_items = (long long)W * CV;
  const long long total = (long long)H * row_items;
  const int seg0 = blockIdx.x * EW_THREADS;
  const int h0 = blockIdx.y * BLUR_ROWS, h1 = min(H, h0 + BLUR_ROWS);
  const uint4* xn = x + (size_t)n * total;
  // items [s0, s1) of image rows [r0, r1) are copied; tile column of item i is i - (seg0 - CV)
  const long long s0 = max(seg0 - CV, 0), s1 = min((long long)seg0 + EW_THREADS + CV, row_items);
  const int r0 = max(h0 - 1, 0), r1 = min(h1 + 1, H);
  if (threadIdx.x == 0) {
    sm100::mbar_init(bar, 1);
    sm100::fence_barrier_init();
    const uint32_t row_bytes = (uint32_t)(s1 - s0) * 16u;
    sm100::mbar_expect_tx(bar, row_bytes * (uint32_t)(r1 - r0));
    for (int h = r0; h < r1; ++h)
      bulk_load_1d(blur_smem + ((size_t)(h - (h0 - 1)) * pitch + (s0 - (seg0 - CV))) * 16, xn + (long long)h * row_items + s0,
                   row_bytes, bar);
  }
  // border blocks: the halo outside the image is zero (the blur's zero padding); disjoint from what the copies write
  const uint4 Z = make_uint4(0u, 0u, 0u, 0u);
  if (h0 == 0)                                         // row above the image
    for (int i = threadIdx.x; i < pitch; i += EW_THREADS) tile[i] = Z;
  if (h1 == H)                                         // row below the image
    for (int i = threadIdx.x; i < pitch; i += EW_THREADS) tile[(size_t)(H - (h0 - 1)) * pitch + i] = Z;
  if (seg0 == 0)                                       // pixel left of the image
    for (int i = threadIdx.x; i < (BLUR_ROWS + 2) * CV; i += EW_THREADS) tile[(size_t)(i >> cv_shift) * pitch + (i & (CV - 1))] = Z;
  const int c1 = (int)(s1 - (seg0 - CV));              // first tile column the copies do not reach
  if (c1 < pitch) {                                    // pixel right of the image (and the tail of a ragged last segment)
    const int nz = pitch - c1;
    for (int i = threadIdx.x; i < (BLUR_ROWS + 2) * nz; i += EW_THREADS) {
      const int r = i / nz;
      tile[(size_t)r * pitch + c1 + (i - r * nz)] = Z;
    }
  }
  if (has_stats)
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int ci = seg0 + threadIdx.x;   // (w, cv) inside a row
  const int cv = ci & (CV - 1);
  float s1a[8], s2a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1a[j] = 0.f; s2a[j] = 0.f; }
  float nw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) nw[j] = has_noise ? noise_w[cv * 8 + j] : 0.f;
  const uint2 nkey = noise_key(seed + (seed_dev ? *seed_dev : 0ull), subseq);
  sm100::mbar_wait(bar, 0);
  if (ci < row_items) {
    // running pointers / indices of this thread's column: one 64-bit add per row instead of a multiply chain
    uint4* yp = y + (size_t)n * total + (long long)h0 * row_items + ci;
    const float* zp0 = noise + ((size_t)n * total + (long long)h0 * row_items + ci) * 8;    // used only when noise != nullptr
    unsigned long long pair0 = ((unsigned long long)n * total + (unsigned long long)h0 * row_items + ci) * 4ull;   // channel 0's pair
    const uint4* tp = tile + CV + threadIdx.x;
    float prev[8], cur[8], nxt[8];
    blur_hrow(tp, CV, prev);
    blur_hrow(tp + pitch, CV, cur);
#pragma unroll
    for (int r = 0; r < BLUR_ROWS; ++r) {
      // straight-line over the 8 rows (the rotation below is register renaming); rows past a ragged bottom tile compute
      // on whatever the unstaged tile rows hold and are dropped at the store / statistics
      const bool live = h0 + r < h1;
      blur_hrow(tp + (r + 2) * pitch, CV, nxt);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (prev[j] + 2.f * cur[j] + nxt[j]) * (1.f / 16.f);
      if (has_noise) {
        float z[8];
        if (!use_rng) {
          float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f), z1 = z0;
          if (live) { z0 = reinterpret_cast<const float4*>(zp0)[0]; z1 = reinterpret_cast<const float4*>(zp0)[1]; }
          z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
        } else {
          normal_oct(nkey, pair0, z);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(nw[j], z[j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = act_fn(acc[j], actv, slope);
      if (live) {
        if (has_stats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s1a[j] += acc[j]; s2a[j] += acc[j] * acc[j]; }
        }
        *yp = pack8(acc);
      }
      yp += row_items; zp0 += row_items * 8; pair0 += (unsigned long long)row_items * 4ull;
#pragma unroll
      for (int j = 0; j < 8; ++j) { prev[j] = cur[j]; cur[j] = nxt[j]; }
    }
  }
  if (has_stats) {
    // lanes congruent mod CV hold the same channels: fold them, then one shared atomic per warp
    if (CV <= 32) {
      for (int off = 16; off >= CV; off >>= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1a[j] += __shfl_xor_sync(0xffffffffu, s1a[j], off);
          s2a[j] += __shfl_xor_sync(0xffffffffu, s2a[j], off);
        }
      }
    }
    const int lane = threadIdx.x & 31;
    if (CV > 32 || lane < CV) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sacc[2 * (cv * 8 + j)], s1a[j]);
        atomicAdd(&sacc[2 * (cv * 8 + j) + 1], s2a[j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[(size_t)n * C * 2 + i], sacc[i]);
  }
}

// ---- generator output: AdaIN apply + 1x1 conv + tanh -------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
gen_output_kernel(const uint4* __restrict__ x, const float* __restrict__ coef, const float* __restrict__ w,
                  const float* __restrict__ b0, long long HW, int C, float* __restrict__ out) {
  extern __shared__ float wa[];  // [C] w*a, then [1] constant
  const int n = blockIdx.y, CV = C / 8;
  if (threadIdx.x < 32) {
    float cst = 0.f;
    for (int c = threadIdx.x; c < C; c += 32) {
      const float a = coef[((size_t)n * C + c) * 2], b = coef[((size_t)n * C + c) * 2 + 1];
      wa[c] = w[c] * a;
      cst += w[c] * b;
    }
    cst = warp_sum(cst);
    if (threadIdx.x == 0) wa[C] = cst + b0[0];
  }
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const uint4* xp = x + ((size_t)n * HW + pix) * CV;
  float acc = wa[C];
  for (int v = 0; v < CV; ++v) {
    float f[8];
    unpack8(xp[v], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(wa[v * 8 + j], f[j], acc);
  }
  out[(size_t)n * HW + pix] = tanhf(acc);
}

// ---- recognizer stem: conv 1->Cout 3x3 pad 1 + ReLU + maxpool 2x2 --------------------------------
__global__ void __launch_bounds__(EW_THREADS)
hwr_stem_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b, int N,
                int H, int W, int Cout, uint4* __restrict__ y) {
  extern __shared__ float ws[];  // [Cout][9] then [Cout]
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[Cout * 9 + i] = b[i];
  __syncthreads();
  const int CV = Cout / 8, Hp = H / 2, Wp = W / 2;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)N * Hp * Wp * CV) return;
  int cv, wp, hp, n;
  decode4(item, CV, Wp, Hp, cv, wp, hp, n);
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
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* k = ws + (cv * 8 + j) * 9;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
      for (int ox = 0; ox < 2; ++ox) {
        float a = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) a = fmaf(k[ky * 3 + kx], patch[oy + ky][ox + kx], a);
        m = fmaxf(m, a);
      }
    f[j] = fmaxf(m + ws[Cout * 9 + cv * 8 + j], 0.f);
  }
  y[item] = pack8(f);
}

__global__ void __launch_bounds__(EW_THREADS)
maxpool_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int C, int kh, int kw,
                    int sh, int sw, int ph, int pw, int Ho, int Wo) {
  const int CV = C / 8;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)N * Ho * Wo * CV) return;
  int cv, wo, ho, n;
  decode4(item, CV, Wo, Ho, cv, wo, ho, n);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -CUDART_INF_F;
  for (int i = 0; i < kh; ++i) {
    const int hh = ho * sh - ph + i;
    if (hh < 0 || hh >= H) continue;
    for (int j2 = 0; j2 < kw; ++j2) {
      const int ww = wo * sw - pw + j2;
      if (ww < 0 || ww >= W) continue;
      float f[8];
      unpack8(x[(((size_t)n * H + hh) * W + ww) * CV + cv], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    }
  }
  y[item] = pack8(m);
}

// 2x2 windows (both poolings of CNNOnlyHWR): the four loads are issued together and the maximum is taken on packed
// bf16 pairs (comparisons of bf16 values are exact) — no unpacking, no dependent load chain
__global__ void __launch_bounds__(EW_THREADS)
maxpool22_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int C, int sh, int sw,
                      int ph, int pw, int Ho, int Wo) {
  const int CV = C / 8;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)N * Ho * Wo * CV) return;
  int cv, wo, ho, n;
  decode4(item, CV, Wo, Ho, cv, wo, ho, n);
  const uint4 NEG = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);   // -inf: padding never wins
  const int h0 = ho * sh - ph, w0 = wo * sw - pw;
  const uint4* xn = x + (size_t)n * H * W * CV;
  uint4 v[4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int hh = h0 + i, ww = w0 + j;
      v[2 * i + j] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? xn[((size_t)hh * W + ww) * CV + cv] : NEG;
    }
  auto mx = [](uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const __nv_bfloat162 m = __hmax2(__hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b)),
                                     __hmax2(*reinterpret_cast<__nv_bfloat162*>(&c), *reinterpret_cast<__nv_bfloat162*>(&d)));
    return *reinterpret_cast<const uint32_t*>(&m);
  };
  y[item] = make_uint4(mx(v[0].x, v[1].x, v[2].x, v[3].x), mx(v[0].y, v[1].y, v[2].y, v[3].y),
                       mx(v[0].z, v[1].z, v[2].z, v[3].z), mx(v[0].w, v[1].w, v[2].w, v[3].w));
}

static inline unsigned blocks_for(long long items, int per_block) {
  return (unsigned)((items + per_block - 1) / per_block);
}
static inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace hwg

using namespace hwg;

extern "C" int hwg_linear_f32(const float* in, const float* W, const float* bias, float* out, int B, int K,
                              int O, int act, float slope, void* stream) {
  HWG_REQUIRE(in && W && out && B > 0 && K > 0 && O > 0, "hwg_linear_f32: bad argument");
  const long long warps = (long long)B * O;
  linear_f32_kernel<<<blocks_for(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(in, W, bias, out, B, K, O, act, slope);
  return check_launch("linear_f32_kernel");
}

extern "C" int hwg_pixelnorm_f32(const float* in, float* out, int B, int K, void* stream) {
  HWG_REQUIRE(in && out && B > 0 && K > 0, "hwg_pixelnorm_f32: bad argument");
  pixelnorm_kernel<<<blocks_for((long long)B * 32, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, K);
  return check_launch("pixelnorm_kernel");
}

extern "C" int hwg_gen_pack_input(const float* content, int64_t cs_t, int64_t cs_b, int64_t cs_c,
                                  const float* style, int T, int B, int C, int S, int Cp, void* x,
                                  void* stream) {
  HWG_REQUIRE(content && x && T > 0 && B > 0 && C > 0 && S >= 0, "hwg_gen_pack_input: bad argument");
  HWG_REQUIRE(S == 0 || style, "hwg_gen_pack_input: null style");
  HWG_REQUIRE(Cp % 8 == 0 && Cp >= C + S, "hwg_gen_pack_input: Cp=%d too small for C+S=%d", Cp, C + S);
  const long long items = (long long)B * T * (Cp / 8);
  gen_pack_input_kernel<<<blocks_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
      content, cs_t, cs_b, cs_c, style, T, B, C, S, Cp, reinterpret_cast<uint4*>(x));
  return check_launch("gen_pack_input_kernel");
}

extern "C" int hwg_adain_coeffs(const float* stats, const float* gamma, const float* beta,
                                int64_t gb_stride_n, int N, int C, int HW, float eps, float* coef,
                                float* save_mean_rstd, void* stream) {
  HWG_REQUIRE(stats && gamma && beta && coef && N > 0 && C > 0 && HW > 0, "hwg_adain_coeffs: bad argument");
  adain_coeffs_kernel<<<blocks_for((long long)N * C, 256), 256, 0, (cudaStream_t)stream>>>(
      stats, gamma, beta, gb_stride_n, N, C, 1.0f / (float)HW, eps, coef, save_mean_rstd);
  return check_launch("adain_coeffs_kernel");
}

extern "C" int hwg_bn_coeffs(const float* stats, int N, int C, int64_t count_per_n, const float* weight,
                             const float* bias, float* running_mean, float* running_var, float momentum,
                             float eps, int use_batch_stats, float* coef, float* save_mean_rstd,
                             void* stream) {
  HWG_REQUIRE(coef && C > 0, "hwg_bn_coeffs: bad argument");
  HWG_REQUIRE(use_batch_stats ? (stats && N > 0 && count_per_n > 0) : (running_mean && running_var),
              "hwg_bn_coeffs: missing statistics");
  bn_coeffs_kernel<<<blocks_for(C, 128), 128, 0, (cudaStream_t)stream>>>(
      stats, N, C, (float)((double)N * (double)count_per_n), weight, bias, running_mean, running_var, momentum,
      eps, use_batch_stats, coef, save_mean_rstd);
  return check_launch("bn_coeffs_kernel");
}

extern "C" int hwg_scale_shift_act(const void* x, void* y, const float* coef, int per_sample, int N,
                                   int64_t HW, int C, int act, float slope, void* stream) {
  HWG_REQUIRE(x && y && coef && N > 0 && HW > 0 && C > 0 && C % 8 == 0, "hwg_scale_shift_act: bad argument");
  const long long total = HW * (C / 8);
  if (pow2(C / 8) && C / 8 <= ST_THREADS) {
    // in-place (x == y) is fine: a block only ever writes items of chunks it has already pulled into its ring
    const long long need = blocks_for(total, ST_CHUNK), cap = (148LL * 3 + N - 1) / N;
    dim3 grid((unsigned)(need < cap ? need : cap), N);
    HWG_SMEM_OPTIN(scale_shift_act_stream_kernel);
    scale_shift_act_stream_kernel<<<grid, ST_THREADS, stream_smem_bytes(1) + ST_BAR_SLOTS * 8 + (size_t)2 * C * sizeof(float),
                                    (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x),
                                                            reinterpret_cast<uint4*>(y), coef, per_sample, HW, C, act, slope);
    return check_launch("scale_shift_act_stream_kernel");
  }
  const long long need = blocks_for(total, EW_THREADS * EW_ITER), cap = (148LL * 8 + N - 1) / N;   // persistent
  dim3 grid((unsigned)(need < cap ? need : cap), N);
  scale_shift_act_kernel<<<grid, EW_THREADS, (size_t)C * 2 * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), coef, per_sample, HW, C, act, slope);
  return check_launch("scale_shift_act_kernel");
}

extern "C" int hwg_blur_noise_act_stats(const void* x, void* y, int N, int H, int W, int C,
                                        const float* noise, const float* noise_w, uint64_t noise_seed,
                                        uint64_t noise_subseq, const uint64_t* noise_seed_dev, int act, float slope,
                                        float* stats, void* stream) {
  HWG_REQUIRE(x && y && x != y && N > 0 && H > 0 && W > 0, "hwg_blur_noise_act_stats: bad argument");
  HWG_REQUIRE(C % 8 == 0 && pow2(C / 8) && C / 8 <= EW_THREADS,
              "hwg_blur_noise_act_stats: C=%d must be 8 x a power of two", C);
  HWG_REQUIRE(noise == nullptr || noise_w != nullptr, "hwg_blur_noise_act_stats: noise needs noise_w");
  const int CV = C / 8;
  int cv_shift = 0;
  while ((1 << cv_shift) < CV) ++cv_shift;
  dim3 grid(blocks_for((long long)W * CV, EW_THREADS), (H + BLUR_ROWS - 1) / BLUR_ROWS, N);
  HWG_REQUIRE(grid.y <= 65535 && N <= 65535, "hwg_blur_noise_act_stats: H=%d / N=%d too large", H, N);
  // the two shapes of the hot path are compile-time specialisations (see the kernel's header)
  auto k = blur_noise_act_stats_kernel<0>;
  if (noise_w && !noise && stats && act == HWG_ACT_LRELU) k = blur_noise_act_stats_kernel<1>;
  else if (!noise_w && !stats && act == HWG_ACT_NONE) k = blur_noise_act_stats_kernel<2>;
  HWG_SMEM_OPTIN(k);
  k<<<grid, EW_THREADS, blur_tile_bytes(CV) + (size_t)C * 2 * sizeof(float) + 16, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), H, W, C, cv_shift, noise, noise_w, noise_seed,
      noise_subseq, reinterpret_cast<const unsigned long long*>(noise_seed_dev), act, slope, stats);
  return check_launch("blur_noise_act_stats_kernel");
}

extern "C" int hwg_gen_output(const void* x, const float* coef, const float* w, const float* b0, int N, int64_t HW,
                              int C, float* out, void* stream) {
  HWG_REQUIRE(x && coef && w && b0 && out && N > 0 && HW > 0 && C > 0 && C % 8 == 0, "hwg_gen_output: bad argument");
  dim3 grid(blocks_for(HW, EW_THREADS), N);
  gen_output_kernel<<<grid, EW_THREADS, (size_t)(C + 1) * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), coef, w, b0, HW, C, out);
  return check_launch("gen_output_kernel");
}

extern "C" int hwg_hwr_stem(const float* img, const float* w, const float* b, int N, int H, int W, int Cout,
                            void* y, void* stream) {
  HWG_REQUIRE(img && w && b && y && N > 0, "hwg_hwr_stem: bad argument");
  HWG_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cout % 8 == 0, "hwg_hwr_stem: H=%d W=%d must be even, Cout=%d a multiple of 8", H, W, Cout);
  const long long items = (long long)N * (H / 2) * (W / 2) * (Cout / 8);
  hwr_stem_kernel<<<blocks_for(items, EW_THREADS), EW_THREADS, (size_t)Cout * 10 * sizeof(float), (cudaStream_t)stream>>>(
      img, w, b, N, H, W, Cout, reinterpret_cast<uint4*>(y));
  return check_launch("hwr_stem_kernel");
}

extern "C" int hwg_maxpool_nhwc(const void* x, void* y, int N, int H, int W, int C, int kh, int kw, int sh,
                                int sw, int ph, int pw, int Ho, int Wo, void* stream) {
  HWG_REQUIRE(x && y && N > 0 && C % 8 == 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0, "hwg_maxpool_nhwc: bad argument");
  HWG_REQUIRE(Ho == (H + 2 * ph - kh) / sh + 1 && Wo == (W + 2 * pw - kw) / sw + 1,
              "hwg_maxpool_nhwc: Ho/Wo do not match the pooling geometry");
  const long long items = (long long)N * Ho * Wo * (C / 8);
  if (kh == 2 && kw == 2) {
    maxpool22_nhwc_kernel<<<blocks_for(items, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), N, H, W, C, sh, sw, ph, pw, Ho, Wo);
    return check_launch("maxpool22_nhwc_kernel");
  }
  maxpool_nhwc_kernel<<<blocks_for(items, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), N, H, W, C, kh, kw, sh, sw, ph, pw, Ho, Wo);
  return check_launch("maxpool_nhwc_kernel");
}
