// Memory-bound passes of the discriminator (reference model/discriminator_ap.py:68-161) around the tensor-core
// convolutions: the 7-wide shift expansion that turns the 1-channel 7x7 input convolution into a 7-tap, 16-channel
// implicit GEMM (and its adjoint back to the image), GroupNorm coefficients and backward, AvgPool2d, and the fused
// backward of [Dropout2d ->] LeakyReLU [-> AvgPool2d].  NHWC bf16, 16-byte vectors of 8 channels, fp32 arithmetic.
// Every kernel's roofline is HBM: activation bytes read + written once per pass.
#include "common.cuh"

namespace hwg {
namespace {

constexpr int DT = 256;      // threads per block
constexpr int SHIFT_C = 16;  // channels of the expanded image (kw <= 16)

__device__ __forceinline__ void unpack8d(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8d(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
inline unsigned grid_for(long long items, int per_block, long long cap = 148LL * 16) {
  long long b = (items + per_block - 1) / per_block;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

// grid of the row-indexed passes: blockIdx.y = image, blockIdx.x strides over the rows (~16 blocks per SM in total)
inline dim3 row_grid(int H, int N) {
  long long bx = H, cap = (148LL * 16 + N - 1) / N;
  if (bx > cap) bx = cap;
  return dim3((unsigned)(bx < 1 ? 1 : bx), (unsigned)N);
}
inline int log2_or_neg(int v) {
  for (int s = 0; s < 31; ++s) if ((1 << s) == v) return s;
  return -1;
}

// ---- image -> [N,H,W,16] bf16, channel j = image[h, w + j - pad] (zero outside), j < kw ----------------------
__global__ void __launch_bounds__(DT)
shift_expand_kernel(const float* __restrict__ img, uint4* __restrict__ out, long long pixels, int W, int kw, int pad) {
  for (long long p = (long long)blockIdx.x * DT + threadIdx.x; p < pixels; p += (long long)gridDim.x * DT) {
    const int w = (int)(p % W);
    const float* row = img + (p - w);
    float f[SHIFT_C];
#pragma unroll
    for (int j = 0; j < SHIFT_C; ++j) {
      const int x = w + j - pad;
      f[j] = (j < kw && x >= 0 && x < W) ? __ldg(row + x) : 0.f;
    }
    float lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = f[j]; hi[j] = f[8 + j]; }
    out[2 * p] = pack8d(lo);
    out[2 * p + 1] = pack8d(hi);
  }
}

// ---- adjoint: dimg[h, x] (+)= sum_j g[h, x - j + pad, j] --------------------------------------------------------
__global__ void __launch_bounds__(DT)
shift_collapse_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ dimg, long long pixels, int W, int kw,
                      int pad, int accumulate) {
  for (long long p = (long long)blockIdx.x * DT + threadIdx.x; p < pixels; p += (long long)gridDim.x * DT) {
    const int x = (int)(p % W);
    const __nv_bfloat16* row = g + (p - x) * SHIFT_C;
    float acc = 0.f;
    for (int j = 0; j < kw; ++j) {
      const int w = x - j + pad;
      if (w >= 0 && w < W) acc += __bfloat162float(row[(long long)w * SHIFT_C + j]);
    }
    dimg[p] = accumulate ? dimg[p] + acc : acc;
  }
}

// ---- GroupNorm coefficients from per-(n,c) sums ---------------------------------------------------------------
// coef[n,c] = (a, b): GroupNorm(z) = a*z + b;  save[n,c] = (mean, rstd) of c's group
__global__ void gn_coeffs_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int C, int cg, float count, float eps,
                                 float* __restrict__ coef, float* __restrict__ save) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / cg) * cg;
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cg; ++k) {
      s1 += stats[((size_t)n * C + g0 + k) * 2];
      s2 += stats[((size_t)n * C + g0 + k) * 2 + 1];
    }
    const float mean = s1 / count;
    const float var = fmaxf(s2 / count - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float a = (gamma ? gamma[c] : 1.f) * rstd;
    coef[((size_t)n * C + c) * 2] = a;
    coef[((size_t)n * C + c) * 2 + 1] = (beta ? beta[c] : 0.f) - mean * a;
    if (save) { save[((size_t)n * C + c) * 2] = mean; save[((size_t)n * C + c) * 2 + 1] = rstd; }
  }
}

// ---- AvgPool2d(kh,kw) (stride = kernel, floor) ----------------------------------------------------------------
__global__ void __launch_bounds__(DT)
avgpool_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int CV, int kh, int kw, int Ho,
               int Wo) {
  const long long total = (long long)N * Ho * Wo * CV;
  const float inv = 1.f / (float)(kh * kw);
  for (long long i = (long long)blockIdx.x * DT + threadIdx.x; i < total; i += (long long)gridDim.x * DT) {
    int cv, wo, ho, n;
    if (i < (1LL << 31)) {     // 32-bit divisions when the index fits
      const unsigned u = (unsigned)i, q1 = u / (unsigned)CV, q2 = q1 / (unsigned)Wo, q3 = q2 / (unsigned)Ho;
      cv = (int)(u - q1 * (unsigned)CV); wo = (int)(q1 - q2 * (unsigned)Wo); ho = (int)(q2 - q3 * (unsigned)Ho); n = (int)q3;
    } else {
      const long long q1 = i / CV, q2 = q1 / Wo, q3 = q2 / Ho;
      cv = (int)(i - q1 * CV); wo = (int)(q1 - q2 * Wo); ho = (int)(q2 - q3 * Ho); n = (int)q3;
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) {
        float f[8];
        unpack8d(x[(((long long)n * H + ho * kh + a) * W + wo * kw + b) * CV + cv], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv;
    y[i] = pack8d(acc);
  }
}

// ---- backward of  y = LeakyReLU(scale[n,c] * conv) [-> AvgPool(kh,kw)] ------------------------------------------
// gz[n,h,w,c] = scale[n,c] * (y > 0 ? 1 : slope) * g[n, h/kh, w/kw, c] / (kh*kw)   (0 where the pool's floor cut)
// Indexing of the three backward passes below (round 2): blockIdx.y = image, blocks stride over the image ROWS, threads
// over the W*CV 16-byte items of a row — all 32-bit; round 1 decoded a flat 64-bit item index with six 64-bit divisions
// per item (~240 instructions around a 16-byte load: ncu showed these passes at 44-56 % issue activity, 3-4 TB/s).
__device__ __forceinline__ int pool_index(int v, int k) { return k == 1 ? v : (k == 2 ? v >> 1 : v / k); }

__global__ void __launch_bounds__(DT)
act_bwd_kernel(const uint4* __restrict__ g, const uint4* __restrict__ y, const float* __restrict__ scale, float slope,
               int H, int W, int CV, int cv_shift, int kh, int kw, int Ho, int Wo, uint4* __restrict__ gz) {
  const int n = blockIdx.y, row_items = W * CV;
  const float inv = 1.f / (float)(kh * kw);
  for (int h = blockIdx.x; h < H; h += gridDim.x) {
    const int ho = pool_index(h, kh);
    const size_t row = ((size_t)n * H + h) * row_items;
    const uint4* grow = g + ((size_t)n * Ho + (ho < Ho ? ho : 0)) * Wo * CV;
    for (int j = threadIdx.x; j < row_items; j += DT) {
      const int w = cv_shift >= 0 ? j >> cv_shift : j / CV, cv = j - w * CV;
      const int wo = pool_index(w, kw);
      float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ho < Ho && wo < Wo) {
        float gv[8], yv[8], sv[8];
        unpack8d(grow[(size_t)wo * CV + cv], gv);
        unpack8d(y[row + j], yv);
        if (scale) {
          const float4* sp = reinterpret_cast<const float4*>(scale + ((size_t)n * CV + cv) * 8);
          const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
          sv[0] = s0.x; sv[1] = s0.y; sv[2] = s0.z; sv[3] = s0.w; sv[4] = s1.x; sv[5] = s1.y; sv[6] = s1.z; sv[7] = s1.w;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) sv[k] = 1.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = gv[k] * inv * sv[k] * (yv[k] > 0.f ? 1.f : slope);
      }
      gz[row + j] = pack8d(o);
    }
  }
}

// ---- backward of  a = LeakyReLU(coef_a*z + coef_b) [-> AvgPool(kh,kw)]  with per-(n,c) coefficients, pass 1 ------
// sums[n,c] += (sum gy', sum gy'*z),  gy' = g[n,h/kh,w/kw,c]/(kh*kw) * (coef_a*z + coef_b > 0 ? 1 : slope)
template <int CV>
__global__ void __launch_bounds__(DT)
norm_bwd_reduce_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const float* __restrict__ coef,
                       float slope, int H, int W, int kh, int kw, int Ho, int Wo, float* __restrict__ sums) {
  constexpr int LANES = DT / CV;                   // pixels in flight per block
  __shared__ float red[LANES][CV * 8 * 2 + 1];
  const int n = blockIdx.y, cv = threadIdx.x % CV, lane = threadIdx.x / CV;
  const int C = CV * 8;
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = coef[((size_t)n * C + cv * 8 + k) * 2];
    b[k] = coef[((size_t)n * C + cv * 8 + k) * 2 + 1];
  }
  float s0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, s1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float inv = 1.f / (float)(kh * kw);
  for (int h = blockIdx.x; h < H; h += gridDim.x) {
    const int ho = pool_index(h, kh);
    if (ho >= Ho) continue;
    const uint4* grow = g + ((size_t)n * Ho + ho) * Wo * CV + cv;
    const uint4* zrow = z + ((size_t)n * H + h) * W * CV + cv;
    for (int w = lane; w < W; w += LANES) {
      const int wo = pool_index(w, kw);
      if (wo >= Wo) continue;
      float gv[8], zv[8];
      unpack8d(grow[(size_t)wo * CV], gv);
      unpack8d(zrow[(size_t)w * CV], zv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gy = gv[k] * inv * (fmaf(a[k], zv[k], b[k]) > 0.f ? 1.f : slope);
        s0[k] += gy;
        s1[k] = fmaf(gy, zv[k], s1[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[lane][(cv * 8 + k) * 2] = s0[k];
    red[lane][(cv * 8 + k) * 2 + 1] = s1[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += DT) {
    float t = 0.f;
#pragma unroll 4
    for (int l = 0; l < LANES; ++l) t += red[l][i];
    atomicAdd(sums + (size_t)n * C * 2 + i, t);
  }
}

// ---- GroupNorm backward coefficients: per-(n,c) (sc, P, Q) with  gz = sc*gy' + P*z + Q --------------------------
// group G of c (cg channels, M = cg*HW elements), r = rstd, mu = mean:
//   T_c = r*(S1_c - mu*S0_c) = sum gy'*xhat;  A = sum_G gamma*S0 / M;  B = sum_G gamma*T / M
//   sc = r*gamma_c;  P = -r*r*B;  Q = -r*A + r*r*mu*B;   dgamma_c += T_c, dbeta_c += S0_c
__global__ void gn_bwd_coeffs_kernel(const float* __restrict__ sums, const float* __restrict__ save,
                                     const float* __restrict__ gamma, int C, int cg, float count,
                                     float* __restrict__ spq, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / cg) * cg;
    const float mu = save[((size_t)n * C + c) * 2], r = save[((size_t)n * C + c) * 2 + 1];
    float A = 0.f, B = 0.f;
    for (int k = 0; k < cg; ++k) {
      const float S0 = sums[((size_t)n * C + g0 + k) * 2], S1 = sums[((size_t)n * C + g0 + k) * 2 + 1];
      const float gm = gamma ? gamma[g0 + k] : 1.f;
      A += gm * S0;
      B += gm * r * (S1 - mu * S0);
    }
    A /= count;
    B /= count;
    spq[((size_t)n * C + c) * 3] = r * (gamma ? gamma[c] : 1.f);
    spq[((size_t)n * C + c) * 3 + 1] = -r * r * B;
    spq[((size_t)n * C + c) * 3 + 2] = -r * A + r * r * mu * B;
    const float S0 = sums[((size_t)n * C + c) * 2], S1 = sums[((size_t)n * C + c) * 2 + 1];
    if (dgamma) atomicAdd(dgamma + c, r * (S1 - mu * S0));
    if (dbeta) atomicAdd(dbeta + c, S0);
  }
}

// ---- pass 2: gz = sc*gy' + P*z + Q ------------------------------------------------------------------------------
__global__ void __launch_bounds__(DT)
norm_bwd_apply_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const float* __restrict__ coef,
                      const float* __restrict__ spq, float slope, int H, int W, int CV, int cv_shift, int kh, int kw,
                      int Ho, int Wo, uint4* __restrict__ gz) {
  // per-channel constants a, b, sc, P, Q as [5][2 halves][C/2] floats: a lane's two float4 reads per constant are
  // 16 bytes apart from its neighbour's (conflict-free); round 1's [C][5] scalar layout was 40 bank-conflicted loads per item
  extern __shared__ __align__(16) float cs[];
  const int n = blockIdx.y, C = CV * 8, C2 = C / 2;
  for (int c = threadIdx.x; c < C; c += DT) {
    const int at = (c & 4 ? C2 : 0) + (c >> 3) * 4 + (c & 3);
    cs[at] = coef[((size_t)n * C + c) * 2];
    cs[C + at] = coef[((size_t)n * C + c) * 2 + 1];
    cs[2 * C + at] = spq[((size_t)n * C + c) * 3];
    cs[3 * C + at] = spq[((size_t)n * C + c) * 3 + 1];
    cs[4 * C + at] = spq[((size_t)n * C + c) * 3 + 2];
  }
  __syncthreads();
  const float inv = 1.f / (float)(kh * kw);
  const int row_items = W * CV;
  for (int h = blockIdx.x; h < H; h += gridDim.x) {
    const int ho = pool_index(h, kh);
    const size_t row = ((size_t)n * H + h) * row_items;
    const uint4* grow = g + ((size_t)n * Ho + (ho < Ho ? ho : 0)) * Wo * CV;
    for (int j = threadIdx.x; j < row_items; j += DT) {
      const int w = cv_shift >= 0 ? j >> cv_shift : j / CV, cv = j - w * CV;
      const int wo = pool_index(w, kw);
      float zv[8], gv[8], o[8];
      unpack8d(z[row + j], zv);
      const bool in = ho < Ho && wo < Wo;
      if (in) unpack8d(grow[(size_t)wo * CV + cv], gv);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const float* base = cs + hf * C2 + cv * 4;
        const float4 ca = *reinterpret_cast<const float4*>(base), cb = *reinterpret_cast<const float4*>(base + C);
        const float4 sc = *reinterpret_cast<const float4*>(base + 2 * C), P = *reinterpret_cast<const float4*>(base + 3 * C);
        const float4 Q = *reinterpret_cast<const float4*>(base + 4 * C);
        const float av[4] = {ca.x, ca.y, ca.z, ca.w}, bv[4] = {cb.x, cb.y, cb.z, cb.w}, sv[4] = {sc.x, sc.y, sc.z, sc.w};
        const float pv[4] = {P.x, P.y, P.z, P.w}, qv[4] = {Q.x, Q.y, Q.z, Q.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = 4 * hf + kk;
          const float gy = in ? gv[k] * inv * (fmaf(av[kk], zv[k], bv[kk]) > 0.f ? 1.f : slope) : 0.f;
          o[k] = fmaf(sv[kk], gy, fmaf(pv[kk], zv[k], qv[kk]));
        }
      }
      gz[row + j] = pack8d(o);
    }
  }
}

// ---- SpectralNorm (discriminator_ap.py:19-32): one power iteration per layer, all layers per launch ------------
// v = normalize(W^T u); u = normalize(W v); inv_sigma = 1 / (u . W v); u, v updated in place.  Three small launches
// over (layer, chunk) grids — a single block per layer spent 226 us on the 256 x 2304 layer:
//   wtu:    v_raw = W^T u        (thread per column),  norms[l][0] += |v_raw|^2
//   wv:     u_raw = W v_raw      (warp per row),       norms[l][1] += |u_raw|^2
//   finish: v = v_raw/|v_raw|, W v = u_raw/|v_raw|, u = W v/|W v|, inv_sigma = |Wv|_eps / |Wv|^2; norms reset to 0
struct SnJob { const float* w; float* u; float* v; int h, wd; };
constexpr int SN_ROWS = 8;                         // rows (warps) per block in the W v pass

constexpr int SN_COLS = 32, SN_RG = DT / SN_COLS;   // a block owns 32 columns; 8 row groups split the rows of a column
__global__ void __launch_bounds__(DT) sn_wtu_kernel(const SnJob* __restrict__ jobs, float* __restrict__ norms) {
  const SnJob j = jobs[blockIdx.y];
  if ((int)blockIdx.x * SN_COLS >= j.wd) return;
  __shared__ float us[1024];
  __shared__ float part[SN_RG][SN_COLS + 1];
  const int cx = threadIdx.x % SN_COLS, rg = threadIdx.x / SN_COLS;
  const int c = blockIdx.x * SN_COLS + cx;
  float acc = 0.f;
  for (int r0 = 0; r0 < j.h; r0 += 1024) {
    const int nr = min(1024, j.h - r0);
    __syncthreads();
    for (int r = threadIdx.x; r < nr; r += DT) us[r] = j.u[r0 + r];
    __syncthreads();
    if (c < j.wd) {
      const float* wc = j.w + (size_t)r0 * j.wd + c;
#pragma unroll 8
      for (int r = rg; r < nr; r += SN_RG) acc = fmaf(__ldg(wc + (size_t)r * j.wd), us[r], acc);
    }
  }
  part[rg][cx] = acc;
  __syncthreads();
  if (rg == 0) {                                   // warp 0: one lane per column
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < SN_RG; ++g) t += part[g][cx];
    float sq = 0.f;
    if (c < j.wd) { j.v[c] = t; sq = t * t; }
    sq = warp_sum(sq);
    if (cx == 0) atomicAdd(norms + 2 * blockIdx.y, sq);
  }
}

__global__ void __launch_bounds__(SN_ROWS * 32) sn_wv_kernel(const SnJob* __restrict__ jobs, float* __restrict__ norms) {
  const SnJob j = jobs[blockIdx.y];
  const int r = blockIdx.x * SN_ROWS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= j.h) return;
  const float* wr = j.w + (size_t)r * j.wd;
  float acc = 0.f;
  for (int c = lane; c < j.wd; c += 32) acc = fmaf(__ldg(wr + c), j.v[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) { j.u[r] = acc; atomicAdd(norms + 2 * blockIdx.y + 1, acc * acc); }
}

__global__ void __launch_bounds__(DT) sn_finish_kernel(const SnJob* __restrict__ jobs, float* __restrict__ norms,
                                                       float* __restrict__ inv_sigma) {
  const SnJob j = jobs[blockIdx.x];
  const float v2 = norms[2 * blockIdx.x], u2 = norms[2 * blockIdx.x + 1];
  const float vn = sqrtf(v2) + 1e-12f;                 // l2normalize: x / (|x| + eps)
  const float wv2 = u2 / (vn * vn);                    // |W v|^2 with the normalised v
  const float un = sqrtf(wv2) + 1e-12f;
  for (int c = threadIdx.x; c < j.wd; c += DT) j.v[c] /= vn;
  for (int r = threadIdx.x; r < j.h; r += DT) j.u[r] = j.u[r] / vn / un;
  __syncthreads();
  if (threadIdx.x == 0) {
    inv_sigma[blockIdx.x] = un / wv2;                  // sigma = u . (W v) = |W v|^2 / (|W v| + eps)
    norms[2 * blockIdx.x] = 0.f;
    norms[2 * blockIdx.x + 1] = 0.f;
  }
}

// ---- per-channel sum of an NHWC bf16 tensor (bias gradients of the discriminator's convolutions) ------------------
template <int CV>
__global__ void __launch_bounds__(DT)
channel_sum_kernel(const uint4* __restrict__ x, long long rows, float* __restrict__ out) {
  constexpr int LANES = DT / CV;
  __shared__ float red[LANES][CV * 8 + 1];
  const int cv = threadIdx.x % CV, lane = threadIdx.x / CV;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long r = (long long)blockIdx.x * LANES + lane; r < rows; r += (long long)gridDim.x * LANES) {
    float f[8];
    unpack8d(x[r * CV + cv], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += f[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[lane][cv * 8 + k] = s[k];
  __syncthreads();
  for (int i = threadIdx.x; i < CV * 8; i += DT) {
    float t = 0.f;
#pragma unroll 4
    for (int l = 0; l < LANES; ++l) t += red[l][i];
    atomicAdd(out + i, t);
  }
}

// ---- SpectralNorm backward: W_eff = W / sigma, sigma = u^T W v with u, v constants (discriminator_ap.py:30-32) ---
// given dW1 = (dL/dW_eff) / sigma (the unpacked wgrad, already scaled):  dW = dW1 - (<dW1, W> / sigma) * u v^T
struct SnGradJob { const float* w; float* gw; const float* u; const float* v; int h, wd; };
__global__ void __launch_bounds__(DT) sn_grad_dot_kernel(const SnGradJob* __restrict__ jobs, float* __restrict__ dots) {
  const SnGradJob j = jobs[blockIdx.y];
  const long long n = (long long)j.h * j.wd;
  __shared__ float red[DT / 32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * DT + threadIdx.x; i < n; i += (long long)gridDim.x * DT)
    acc = fmaf(j.gw[i], j.w[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < DT / 32; ++i) t += red[i];
    if (t != 0.f) atomicAdd(dots + blockIdx.y, t);
  }
}
__global__ void __launch_bounds__(DT) sn_grad_apply_kernel(const SnGradJob* __restrict__ jobs, const float* __restrict__ dots,
                                                           const float* __restrict__ inv_sigma) {
  const SnGradJob j = jobs[blockIdx.y];
  const long long n = (long long)j.h * j.wd;
  const float k = dots[blockIdx.y] * inv_sigma[blockIdx.y];
  for (long long i = (long long)blockIdx.x * DT + threadIdx.x; i < n; i += (long long)gridDim.x * DT) {
    const int r = (int)(i / j.wd), c = (int)(i % j.wd);
    j.gw[i] -= k * j.u[r] * j.v[c];
  }
}

bool cv_supported(int C) { return C == 16 || C == 32 || C == 64 || C == 128 || C == 256; }

}  // namespace
}  // namespace hwg

using namespace hwg;

extern "C" int hwg_shift_expand(const float* img, void* out, int N, int H, int W, int kw, int pad, void* stream) {
  HWG_REQUIRE(img && out && N > 0 && H > 0 && W > 0 && kw >= 1 && kw <= SHIFT_C && pad >= 0, "hwg_shift_expand: bad argument");
  const long long pixels = (long long)N * H * W;
  shift_expand_kernel<<<grid_for(pixels, DT), DT, 0, (cudaStream_t)stream>>>(img, reinterpret_cast<uint4*>(out), pixels, W,
                                                                            kw, pad);
  return check_launch("shift_expand_kernel");
}

extern "C" int hwg_shift_collapse(const void* g, float* dimg, int N, int H, int W, int kw, int pad, int accumulate,
                                  void* stream) {
  HWG_REQUIRE(g && dimg && N > 0 && H > 0 && W > 0 && kw >= 1 && kw <= SHIFT_C && pad >= 0, "hwg_shift_collapse: bad argument");
  const long long pixels = (long long)N * H * W;
  shift_collapse_kernel<<<grid_for(pixels, DT), DT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(g),
                                                                              dimg, pixels, W, kw, pad, accumulate);
  return check_launch("shift_collapse_kernel");
}

extern "C" int hwg_gn_coeffs(const float* stats, const float* gamma, const float* beta, int N, int C, int groups,
                             int64_t HW, float eps, float* coef, float* save_mean_rstd, void* stream) {
  HWG_REQUIRE(stats && coef && N > 0 && C > 0 && groups > 0 && C % groups == 0 && HW > 0, "hwg_gn_coeffs: bad argument");
  const int cg = C / groups;
  gn_coeffs_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(stats, gamma, beta, C, cg, (float)((double)cg * (double)HW), eps,
                                                        coef, save_mean_rstd);
  return check_launch("gn_coeffs_kernel");
}

extern "C" int hwg_avgpool_nhwc(const void* x, void* y, int N, int H, int W, int C, int kh, int kw, void* stream) {
  HWG_REQUIRE(x && y && N > 0 && C > 0 && C % 8 == 0 && kh >= 1 && kw >= 1 && H >= kh && W >= kw, "hwg_avgpool_nhwc: bad argument");
  const int Ho = H / kh, Wo = W / kw;
  avgpool_kernel<<<grid_for((long long)N * Ho * Wo * (C / 8), DT), DT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), N, H, W, C / 8, kh, kw, Ho, Wo);
  return check_launch("avgpool_kernel");
}

extern "C" int hwg_act_bwd(const void* g, const void* y, const float* scale, float slope, int N, int H, int W, int C,
                           int kh, int kw, void* gz, void* stream) {
  HWG_REQUIRE(g && y && gz && N > 0 && C > 0 && C % 8 == 0 && kh >= 1 && kw >= 1 && H >= kh && W >= kw, "hwg_act_bwd: bad argument");
  HWG_REQUIRE((long long)W * (C / 8) < (1LL << 30) && N <= 65535, "hwg_act_bwd: row of %d x %d channels / N=%d too large", W, C, N);
  act_bwd_kernel<<<row_grid(H, N), DT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(y), scale, slope, H, W, C / 8, log2_or_neg(C / 8), kh,
      kw, H / kh, W / kw, reinterpret_cast<uint4*>(gz));
  return check_launch("act_bwd_kernel");
}

extern "C" int hwg_norm_bwd_reduce(const void* g, const void* z, const float* coef, float slope, int N, int H, int W,
                                   int C, int kh, int kw, float* sums, void* stream) {
  HWG_REQUIRE(g && z && coef && sums && N > 0 && kh >= 1 && kw >= 1 && H >= kh && W >= kw, "hwg_norm_bwd_reduce: bad argument");
  HWG_REQUIRE(cv_supported(C), "hwg_norm_bwd_reduce: C=%d not in {16,32,64,128,256}", C);
  HWG_REQUIRE(N <= 65535, "hwg_norm_bwd_reduce: N=%d too large", N);
  const int CV = C / 8;
  const dim3 grid = row_grid(H, N);
  const uint4* gp = reinterpret_cast<const uint4*>(g);
  const uint4* zp = reinterpret_cast<const uint4*>(z);
  const int Ho = H / kh, Wo = W / kw;
  cudaStream_t s = (cudaStream_t)stream;
  switch (CV) {
    case 2: norm_bwd_reduce_kernel<2><<<grid, DT, 0, s>>>(gp, zp, coef, slope, H, W, kh, kw, Ho, Wo, sums); break;
    case 4: norm_bwd_reduce_kernel<4><<<grid, DT, 0, s>>>(gp, zp, coef, slope, H, W, kh, kw, Ho, Wo, sums); break;
    case 8: norm_bwd_reduce_kernel<8><<<grid, DT, 0, s>>>(gp, zp, coef, slope, H, W, kh, kw, Ho, Wo, sums); break;
    case 16: norm_bwd_reduce_kernel<16><<<grid, DT, 0, s>>>(gp, zp, coef, slope, H, W, kh, kw, Ho, Wo, sums); break;
    default: norm_bwd_reduce_kernel<32><<<grid, DT, 0, s>>>(gp, zp, coef, slope, H, W, kh, kw, Ho, Wo, sums); break;
  }
  return check_launch("norm_bwd_reduce_kernel");
}

extern "C" int hwg_gn_bwd_coeffs(const float* sums, const float* save_mean_rstd, const float* gamma, int N, int C,
                                 int groups, int64_t HW, float* spq, float* dgamma, float* dbeta, void* stream) {
  HWG_REQUIRE(sums && save_mean_rstd && spq && N > 0 && C > 0 && groups > 0 && C % groups == 0 && HW > 0,
              "hwg_gn_bwd_coeffs: bad argument");
  const int cg = C / groups;
  gn_bwd_coeffs_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(sums, save_mean_rstd, gamma, C, cg,
                                                            (float)((double)cg * (double)HW), spq, dgamma, dbeta);
  return check_launch("gn_bwd_coeffs_kernel");
}

extern "C" int hwg_norm_bwd_apply(const void* g, const void* z, const float* coef, const float* spq, float slope, int N,
                                  int H, int W, int C, int kh, int kw, void* gz, void* stream) {
  HWG_REQUIRE(g && z && coef && spq && gz && N > 0 && C > 0 && C % 8 == 0 && kh >= 1 && kw >= 1 && H >= kh && W >= kw,
              "hwg_norm_bwd_apply: bad argument");
  HWG_REQUIRE(C <= 1024, "hwg_norm_bwd_apply: C=%d too large", C);
  HWG_REQUIRE((long long)W * (C / 8) < (1LL << 30) && N <= 65535, "hwg_norm_bwd_apply: row of %d x %d channels / N=%d too large", W, C, N);
  norm_bwd_apply_kernel<<<row_grid(H, N), DT, (size_t)C * 5 * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(z), coef, spq, slope, H, W, C / 8, log2_or_neg(C / 8), kh,
      kw, H / kh, W / kw, reinterpret_cast<uint4*>(gz));
  return check_launch("norm_bwd_apply_kernel");
}

extern "C" int hwg_spectral_norm(const void* jobs_dev, int njobs, int max_h, int max_wd, float* norms_scratch,
                                 float* inv_sigma, void* stream) {
  HWG_REQUIRE(jobs_dev && inv_sigma && norms_scratch && njobs > 0 && max_h > 0 && max_wd > 0, "hwg_spectral_norm: bad argument");
  const SnJob* jobs = reinterpret_cast<const SnJob*>(jobs_dev);
  cudaStream_t s = (cudaStream_t)stream;
  sn_wtu_kernel<<<dim3((max_wd + SN_COLS - 1) / SN_COLS, njobs), DT, 0, s>>>(jobs, norms_scratch);
  if (int rc = check_launch("sn_wtu_kernel")) return rc;
  sn_wv_kernel<<<dim3((max_h + SN_ROWS - 1) / SN_ROWS, njobs), SN_ROWS * 32, 0, s>>>(jobs, norms_scratch);
  if (int rc = check_launch("sn_wv_kernel")) return rc;
  sn_finish_kernel<<<njobs, DT, 0, s>>>(jobs, norms_scratch, inv_sigma);
  return check_launch("sn_finish_kernel");
}

extern "C" int hwg_channel_sum(const void* x, int64_t rows, int C, float* out, void* stream) {
  HWG_REQUIRE(x && out && rows > 0, "hwg_channel_sum: bad argument");
  HWG_REQUIRE(cv_supported(C), "hwg_channel_sum: C=%d not in {16,32,64,128,256}", C);
  const int CV = C / 8, lanes = DT / CV;
  long long bx = (rows + lanes - 1) / lanes;
  if (bx > 148LL * 8) bx = 148LL * 8;
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  cudaStream_t s = (cudaStream_t)stream;
  switch (CV) {
    case 2: channel_sum_kernel<2><<<(unsigned)bx, DT, 0, s>>>(xp, rows, out); break;
    case 4: channel_sum_kernel<4><<<(unsigned)bx, DT, 0, s>>>(xp, rows, out); break;
    case 8: channel_sum_kernel<8><<<(unsigned)bx, DT, 0, s>>>(xp, rows, out); break;
    case 16: channel_sum_kernel<16><<<(unsigned)bx, DT, 0, s>>>(xp, rows, out); break;
    default: channel_sum_kernel<32><<<(unsigned)bx, DT, 0, s>>>(xp, rows, out); break;
  }
  return check_launch("channel_sum_kernel");
}

extern "C" int hwg_spectral_norm_bwd(const void* jobs_dev, int njobs, int64_t max_elems, const float* inv_sigma,
                                     float* dots_scratch, void* stream) {
  HWG_REQUIRE(jobs_dev && inv_sigma && dots_scratch && njobs > 0 && max_elems > 0, "hwg_spectral_norm_bwd: bad argument");
  const SnGradJob* jobs = reinterpret_cast<const SnGradJob*>(jobs_dev);
  long long bx = (max_elems + DT * 8 - 1) / (DT * 8);
  if (bx > 64) bx = 64;
  cudaStream_t s = (cudaStream_t)stream;
  sn_grad_dot_kernel<<<dim3((unsigned)bx, njobs), DT, 0, s>>>(jobs, dots_scratch);
  if (int rc = check_launch("sn_grad_dot_kernel")) return rc;
  sn_grad_apply_kernel<<<dim3((unsigned)bx, njobs), DT, 0, s>>>(jobs, dots_scratch, inv_sigma);
  return check_launch("sn_grad_apply_kernel");
}
