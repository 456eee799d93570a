// One-input-channel stem convolutions straight from the fp32 image (sm_100a): the 7x7 `in_conv` of DiscriminatorAP
// (reference model/discriminator_ap.py:75) and the 5x5 `down_conv1[0]` of Encoder2 (model/autoencoder.py:345).
//
// Round 1 ran these as a kh-tap implicit GEMM over a 16-channel shift expansion of the image on the tcgen05 kernel
// (hwg_shift_expand + hwg_conv_fprop).  With K = 16 per tap that launch is all epilogue and 32-byte TMA rows: 861 us at
// 7 % tensor-pipe activity for the discriminator at 128 lines, the slowest convolution launch of the step (ncu, round 2).
// The layer is 4 bytes in and 2*Cout bytes out per pixel, i.e. a WRITE-bound pass, so this kernel
//   * keeps a bf16 copy of the image tile (TR + kh - 1 rows) in shared memory, twice — as is and shifted by one pixel,
//     so that the two horizontally adjacent pixels an mma.sync A register holds are one aligned 32-bit load;
//   * builds the im2col fragments on the fly: K = 8 * kh (kernel row dy -> k = 8*dy + dx, column 7 is zero), one
//     m16n8k16 K step = two kernel rows; the weights (the tap-major [kh][Cout][16] operand the shift-expansion route
//     uses, columns >= kw ignored) live in registers as B fragments for the CTA's whole tile range;
//   * prefetches the next tile's pixels into registers while the current tile is computed (one barrier per tile);
//   * transposes each warp's 16-pixel output tile through a swizzled shared-memory tile, so that global stores are
//     16-byte vectors covering whole 128-byte pixel rows; bias and the per-(n, c) GroupNorm statistics (kept in registers
//     across the tiles of an image) ride on the accumulator fragments.
#include "common.cuh"
#include <string.h>

namespace hwg {
namespace {

constexpr int ST_TR = 8, ST_TC = 64;             // output tile: rows x columns
constexpr int ST_ROWS = ST_TR + 7;               // image rows a tile needs (kh <= 8)
constexpr int ST_PITCH = ST_TC + 8 + 8;          // bf16 elements per tile row (kw - 1 <= 7 halo columns, padded)
constexpr int ST_THREADS = 256;
constexpr int ST_FILL = (ST_ROWS * ST_PITCH + ST_THREADS - 1) / ST_THREADS;   // pixels a thread stages per tile

struct StemParams {
  const float* img; const __nv_bfloat16* w; const float* bias; __nv_bfloat16* y; float* stats;
  int N, H, W, Ho, Wo, kh, kw, pad_h, pad_w;
  int tiles_r, tiles_c, total_tiles;
};

__device__ __forceinline__ void st_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT = Cout / 8.  A warp owns NW = 4 of the NT n-tiles (32 channels): with all 64 channels per warp the B fragments,
// accumulators and statistics took 235 registers = 8 warps per SM; split, two CTAs fit.
template <int NT>
__global__ void __launch_bounds__(ST_THREADS, 2)
stem_conv_kernel(const StemParams p) {
  constexpr int NW = 4, NHALF = NT / NW;
  constexpr int COUT = NT * 8, WPP = NW * 4;     // 32-bit words per output pixel of a warp's channel slice
  __shared__ __align__(16) __nv_bfloat16 tile_e[2][ST_ROWS * ST_PITCH];   // image tile, as is       [buffer][row][col]
  __shared__ __align__(16) __nv_bfloat16 tile_o[2][ST_ROWS * ST_PITCH];   // shifted by one pixel: tile_o[c] = tile_e[c+1]
  __shared__ __align__(16) uint32_t outst[ST_THREADS / 32][16 * WPP];     // per warp: 16 pixels x 32 channels bf16, swizzled
  __shared__ __align__(16) float bias_s[COUT];
  __shared__ float stat_s[2 * COUT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int half = warp % NHALF, ch0 = half * NW * 8;      // this warp's channel slice [ch0, ch0 + 32)

  const int per = (p.total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(p.total_tiles, t_begin + per);
  for (int c = threadIdx.x; c < COUT; c += blockDim.x) bias_s[c] = p.bias ? p.bias[c] : 0.f;
  for (int c = threadIdx.x; c < 2 * COUT; c += blockDim.x) stat_s[c] = 0.f;
  if (t_begin >= t_end) return;

  // B fragments: b[s][nt][0] = w[dy = 2s][co = 8nt + g][dx = 2tq, 2tq+1], [1] the same of kernel row 2s+1
  uint32_t bfr[4][NW][2];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int nt = 0; nt < NW; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int dy = 2 * s + r;
        uint32_t v = 0u;
        if (dy < p.kh && 2 * tq < p.kw) {
          v = *reinterpret_cast<const uint32_t*>(p.w + ((size_t)dy * COUT + ch0 + 8 * nt + g) * 16 + 2 * tq);
          if (2 * tq + 1 >= p.kw) v &= 0xFFFFu;
        }
        bfr[s][nt][r] = v;
      }

  auto tile_coords = [&](int t, int& n, int& r0, int& c0) {
    const int tc = t % p.tiles_c, q = t / p.tiles_c;
    n = q / p.tiles_r; r0 = (q - n * p.tiles_r) * ST_TR; c0 = tc * ST_TC;
  };
  // staging: element e of the tile = (row e / PITCH, column e % PITCH) = image pixel (r0 - pad_h + row, c0 - pad_w + col)
  float pre[ST_FILL];
  auto fetch = [&](int t) {
    int n, r0, c0;
    tile_coords(t, n, r0, c0);
    const float* im = p.img + (size_t)n * p.H * p.W;
#pragma unroll
    for (int i = 0; i < ST_FILL; ++i) {
      const int e = i * ST_THREADS + (int)threadIdx.x;
      const int row = e / ST_PITCH, col = e - row * ST_PITCH;
      const int h = r0 - p.pad_h + row, w = c0 - p.pad_w + col;
      pre[i] = (e < ST_ROWS * ST_PITCH && h >= 0 && h < p.H && w >= 0 && w < p.W) ? __ldg(im + (size_t)h * p.W + w) : 0.f;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < ST_FILL; ++i) {
      const int e = i * ST_THREADS + (int)threadIdx.x;
      if (e < ST_ROWS * ST_PITCH) {
        const __nv_bfloat16 v = __float2bfloat16_rn(pre[i]);
        tile_e[buf][e] = v;
        if (e % ST_PITCH != 0) tile_o[buf][e - 1] = v;
      }
    }
  };

  float s1[NW][2], s2[NW][2];     // statistics of the current image: channels ch0 + 8nt + 2tq, +1
#pragma unroll
  for (int nt = 0; nt < NW; ++nt) { s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f; }
  int stat_n = -1;
  auto flush_stats = [&](int n_img) {
#pragma unroll
    for (int nt = 0; nt < NW; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float u = s1[nt][e], v = s2[nt][e];
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) { u += __shfl_xor_sync(0xffffffffu, u, o); v += __shfl_xor_sync(0xffffffffu, v, o); }
        if (g == 0) {
          atomicAdd(&stat_s[2 * (ch0 + 8 * nt + 2 * tq + e)], u);
          atomicAdd(&stat_s[2 * (ch0 + 8 * nt + 2 * tq + e) + 1], v);
        }
        s1[nt][e] = 0.f; s2[nt][e] = 0.f;
      }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * COUT; c += blockDim.x) {
      atomicAdd(&p.stats[(size_t)n_img * COUT * 2 + c], stat_s[c]);
      stat_s[c] = 0.f;
    }
    __syncthreads();
  };

  fetch(t_begin);
  stage(0);
  __syncthreads();
  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) & 1;
    int n, r0, c0;
    tile_coords(t, n, r0, c0);
    if (p.stats && n != stat_n) {
      if (stat_n >= 0) flush_stats(stat_n);
      stat_n = n;
    }
    if (t + 1 < t_end) fetch(t + 1);            // global loads of the next tile fly while this one is computed

    // columns p and p+1 with p = cs + g + 2tq: parity of p = parity of g -> even: tile_e[p], odd: tile_o[p-1]
    const __nv_bfloat16* src = (g & 1) ? tile_o[buf] - 1 : tile_e[buf];
    uint32_t* ost = outst[warp];
    for (int mt = warp / NHALF; mt < ST_TR * (ST_TC / 16); mt += (ST_THREADS / 32) / NHALF) {
      const int ro = mt / (ST_TC / 16), cs = (mt - ro * (ST_TC / 16)) * 16;
      float acc[NW][4];
#pragma unroll
      for (int nt = 0; nt < NW; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
      const __nv_bfloat16* a_base = src + ro * ST_PITCH + cs + g + 2 * tq;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (2 * s < p.kh) {
          uint32_t a[4];
          const __nv_bfloat16* r_lo = a_base + (2 * s) * ST_PITCH;
          a[0] = *reinterpret_cast<const uint32_t*>(r_lo);
          a[1] = *reinterpret_cast<const uint32_t*>(r_lo + 8);
          a[2] = *reinterpret_cast<const uint32_t*>(r_lo + ST_PITCH);
          a[3] = *reinterpret_cast<const uint32_t*>(r_lo + ST_PITCH + 8);
#pragma unroll
          for (int nt = 0; nt < NW; ++nt) st_mma(acc[nt], a, bfr[s][nt][0], bfr[s][nt][1]);
        }
      }
      // ---- epilogue: rows g / g+8 of the m-tile = pixels (r0 + ro, c0 + cs + g / + 8), channels 8nt + 2tq, +1
      const int ho = r0 + ro;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int px = g + 8 * h;
        const bool valid = ho < p.Ho && c0 + cs + px < p.Wo;
#pragma unroll
        for (int nt = 0; nt < NW; ++nt) {
          const float2 bb = *reinterpret_cast<const float2*>(bias_s + ch0 + 8 * nt + 2 * tq);
          const float v0 = acc[nt][2 * h] + bb.x, v1 = acc[nt][2 * h + 1] + bb.y;
          if (valid) {
            s1[nt][0] += v0; s1[nt][1] += v1;
            s2[nt][0] = fmaf(v0, v0, s2[nt][0]); s2[nt][1] = fmaf(v1, v1, s2[nt][1]);
          }
          const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
          ost[px * WPP + ((4 * nt + tq) ^ ((px & (NW - 1)) << 2))] = *reinterpret_cast<const uint32_t*>(&pk);
        }
      }
      __syncwarp();
      // 16 pixels x NW 16-byte chunks, lanes over (pixel, chunk): 64 contiguous bytes of a pixel row per group of 4 lanes
      if (ho < p.Ho) {
        __nv_bfloat16* yrow = p.y + (((size_t)n * p.Ho + ho) * p.Wo + c0 + cs) * COUT + ch0;
#pragma unroll
        for (int i = 0; i < (16 * NW) / 32; ++i) {
          const int c = i * 32 + lane, px = c / NW, q = c - px * NW;
          if (c0 + cs + px < p.Wo) {
            const uint4 v = *reinterpret_cast<const uint4*>(&ost[px * WPP + ((q ^ (px & (NW - 1))) << 2)]);
            *reinterpret_cast<uint4*>(yrow + (size_t)px * COUT + q * 8) = v;
          }
        }
      }
      __syncwarp();
    }
    if (t + 1 < t_end) stage(buf ^ 1);
    __syncthreads();
  }
  if (p.stats && stat_n >= 0) flush_stats(stat_n);
}

}  // namespace
}  // namespace hwg

using namespace hwg;

extern "C" int hwg_stem_conv(const float* img, const void* w, const float* bias, int N, int H, int W, int kh, int kw,
                             int pad_h, int pad_w, int Cout, void* y, float* stats, void* stream) {
  HWG_REQUIRE(img && w && y && N > 0 && H > 0 && W > 0, "hwg_stem_conv: bad argument");
  HWG_REQUIRE(kh >= 1 && kh <= 8 && kw >= 1 && kw <= 8, "hwg_stem_conv: kernel %dx%d (up to 8x8)", kh, kw);
  HWG_REQUIRE(Cout == 32 || Cout == 64, "hwg_stem_conv: Cout=%d must be 32 or 64", Cout);
  HWG_REQUIRE(pad_h >= 0 && pad_h < kh && pad_w >= 0 && pad_w < kw, "hwg_stem_conv: padding (%d,%d)", pad_h, pad_w);
  HWG_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 3) == 0,
              "hwg_stem_conv: y must be 16-byte aligned");
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.img = img; p.w = reinterpret_cast<const __nv_bfloat16*>(w); p.bias = bias;
  p.y = reinterpret_cast<__nv_bfloat16*>(y); p.stats = stats;
  p.N = N; p.H = H; p.W = W; p.kh = kh; p.kw = kw; p.pad_h = pad_h; p.pad_w = pad_w;
  p.Ho = H + 2 * pad_h - kh + 1; p.Wo = W + 2 * pad_w - kw + 1;
  HWG_REQUIRE(p.Ho > 0 && p.Wo > 0, "hwg_stem_conv: empty output");
  p.tiles_r = (p.Ho + ST_TR - 1) / ST_TR; p.tiles_c = (p.Wo + ST_TC - 1) / ST_TC;
  const long long total = (long long)p.tiles_r * p.tiles_c * N;
  HWG_REQUIRE(total < (1LL << 31), "hwg_stem_conv: too many tiles");
  p.total_tiles = (int)total;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms * 2 < p.total_tiles ? sms * 2 : p.total_tiles;
  if (Cout == 64) stem_conv_kernel<8><<<grid, ST_THREADS, 0, (cudaStream_t)stream>>>(p);
  else stem_conv_kernel<4><<<grid, ST_THREADS, 0, (cudaStream_t)stream>>>(p);
  return check_launch("stem_conv_kernel");
}
