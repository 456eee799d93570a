// Convolution forward / input-gradient for the small-channel layers (Cin in {16,32,64}, <= 64 output channels per
// fold) — the HBM-bound end of the generator (blocks 3-4: 64x1024 and 32x512 images with 16 / 32 channels).
//
// Why not the tcgen05 kernel of hwg_conv.cu: with 16-32 channels a (tap, chunk) operand tile is a TMA box of
// 32-64-byte rows, every tap re-fetches the tile from L2 (9-16x read amplification at the L2 -> SM link, which is
// only ~2x HBM bandwidth on B200) and a 128xN MMA with N = 16 leaves the tensor pipe idle anyway.  These layers
// need bandwidth, not FLOPs, so this kernel stages each input tile ONCE:
//   * persistent CTAs walk TI x TJ output tiles; one TMA box brings the input halo tile (zero-filled outside the
//     image = the reference's zero padding) into a multi-stage shared-memory ring; the packed weights of all taps
//     are loaded once per CTA;
//   * each warp owns 16-pixel row segments: for every tap, ldmatrix reads the staged tile at the tap's pixel
//     offset (stride 2 for the input gradient of the up-sampling convolutions) and mma.sync.m16n8k16 accumulates
//     in registers — the halo tile is read from HBM/L2 once, all re-use happens in shared memory;
//   * the epilogue works on the accumulator fragments: bias, NoiseInjection (counter-based N(0,1) generated per
//     element pair), LeakyReLU, per-(n,c) statistics kept in registers across the CTA's tiles, bf16 stores.
// Up-sampling transposed convolutions run as ONE launch: fold f (= output parity) has its own taps and writes
// its own output pixels (see hwgConvDesc.fold_taps).
#include "common.cuh"
#include "sm100.cuh"
#include "noise_rng.cuh"
#include <cuda.h>
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace hwg {
using namespace sm100;

struct CsParams {
  int N, Ho, Wo, Cout, Cf;     // Cf = channels per fold
  int TI, TJ, tiles_i, tiles_j, total_tiles;
  int sh, sw;                  // input strides
  int dh_min, dw_min, xbox_w;
  int tpf, ntaps;              // taps per fold
  int x_bytes, tx_bytes, stages, w_boxes, w_box_bytes, w_bytes;
  int tap_dh[HWG_MAX_TAPS], tap_dw[HWG_MAX_TAPS];
  long long ysn, ysh, ysw, fold_sh, fold_sw;
  long long zsn, zsh, zsw;
  int fold_w;
  int act, noise_mode, has_stats;
  float slope;
  const float* bias; const float* noise; const float* noise_w; float* stats;
  __nv_bfloat16* y;
  unsigned long long noise_seed, noise_subseq;
  const unsigned long long* noise_seed_dev;
};

__device__ __forceinline__ void cs_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cs_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (row, 16-byte chunk) -> byte offset in a TMA-swizzled tile with PITCH-byte rows (32B / 64B / 128B swizzle)
template <int PITCH>
__device__ __forceinline__ uint32_t cs_swz(uint32_t row, uint32_t chunk) {
  const uint32_t o = row * PITCH + chunk * 16u;
  return o ^ (((o >> 7) & (PITCH / 16u - 1u)) << 4);
}

constexpr int CS_WARPS = 8;

// CIN input channels, NTF 8-channel n-tiles per fold, F folds, MB 16-pixel m-tiles per warp pass, STATS: per-(n,c)
// statistics.  Two CTAs per SM (<= 128 registers): the per-element epilogue is latency-bound, so resident warps
// matter more than the size of a warp's register tile.
// ACT_T / NOISE_T >= 0 fix the epilogue's activation / noise mode at compile time (-1: runtime values of CsParams):
// the epilogue is ~2/3 of this kernel's instructions, and with every option behind a runtime branch the 16 -> 16
// instantiation was 8192 SASS instructions with 23 % of its stall samples on instruction fetch (ncu, round 2).
template <int CIN, int NTF, int F, int MB, int STATS, int ACT_T = -1, int NOISE_T = -1>
__global__ void __launch_bounds__(CS_WARPS * 32, 2)
conv_small_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CsParams p) {
  constexpr int KS = CIN / 16, PX = CIN * 2, NT = NTF * F;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* wsm = smem + (size_t)p.stages * p.x_bytes;
  float* bias_s = reinterpret_cast<float*>(wsm + p.w_bytes);   // [Cout]
  float* nw_s = bias_s + NT * 8;                               // [Cout]
  float* stat_s = nw_s + NT * 8;                               // [Cf][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stat_s + 2 * NTF * 8);
  uint64_t* w_bar = full_bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int per = (p.total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(p.total_tiles, t_begin + per);
  const int ntile = t_end - t_begin;

  auto issue = [&](int it) {
    const int t = t_begin + it, s = it % p.stages;
    const int tj = t % p.tiles_j, r = t / p.tiles_j;
    const int ti = r % p.tiles_i, n = r / p.tiles_i;
    mbar_expect_tx(&full_bar[s], (uint32_t)p.tx_bytes);
    tma_load_4d(smem + (size_t)s * p.x_bytes, &tmap_x, &full_bar[s], 0, tj * p.TJ * p.sw + p.dw_min,
                ti * p.TI * p.sh + p.dh_min, n);
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
    mbar_init(w_bar, 1);
    fence_barrier_init();
    if (ntile > 0) {
      mbar_expect_tx(w_bar, (uint32_t)(p.w_boxes * p.w_box_bytes));
      for (int b = 0; b < p.w_boxes; ++b) tma_load_2d(wsm + (size_t)b * p.w_box_bytes, &tmap_w, w_bar, 0, b * 256);
      for (int it = 0; it < min(p.stages, ntile); ++it) issue(it);
    }
  }
  for (int c = threadIdx.x; c < NT * 8; c += blockDim.x) {
    bias_s[c] = (p.bias && c < p.Cout) ? p.bias[c] : 0.f;
    nw_s[c] = (p.noise_w && c < p.Cout) ? p.noise_w[c] : 0.f;
  }
  for (int c = threadIdx.x; c < 2 * NTF * 8; c += blockDim.x) stat_s[c] = 0.f;
  __syncthreads();
  if (ntile <= 0) return;
  mbar_wait(w_bar, 0);

  // ldmatrix lane roles
  const int mi = lane >> 3, r8 = lane & 7;
  const int a_row = r8 + 8 * (mi & 1), a_kc = mi >> 1;        // A: (rows 0-7,k0) (rows 8-15,k0) (rows 0-7,k1) (rows 8-15,k1)
  const int b_row = r8 + 8 * (mi >> 1), b_kc = mi & 1;        // B: (n 0-7,k0) (n 0-7,k1) (n 8-15,k0) (n 8-15,k1)
  const int g = lane >> 2, tq = lane & 3;
  const uint32_t w_base = smem_u32(wsm);
  const int mt_j = p.TJ >> 4, mt_tile = p.TI * mt_j;
  const uint2 nkey = noise_key(p.noise_seed + (p.noise_seed_dev ? *p.noise_seed_dev : 0ull), p.noise_subseq);
  constexpr bool has_stats = STATS != 0;
  constexpr int NS = STATS ? NT : 1;
  const int act = ACT_T >= 0 ? ACT_T : p.act;
  const int noise_mode = NOISE_T >= 0 ? NOISE_T : p.noise_mode;

  float s1[NS][2], s2[NS][2];   // per-thread statistics of the current image (channels 8*nt + 2tq, +1)
#pragma unroll
  for (int a = 0; a < NS; ++a) { s1[a][0] = s1[a][1] = s2[a][0] = s2[a][1] = 0.f; }
  int stat_n = -1;
  long long fold_off[F];        // element displacement of each fold's output pixel
#pragma unroll
  for (int f = 0; f < F; ++f) fold_off[f] = (f / p.fold_w) * p.fold_sh + (f % p.fold_w) * p.fold_sw;
  const int half_cf = p.Cf >> 1;

  auto flush_stats = [&](int n_img) {
    // fold the 8 pixel rows (g) of the warp, then the warps and folds of the CTA in shared memory
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float u = s1[a][e], v = s2[a][e];
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) { u += __shfl_xor_sync(0xffffffffu, u, o); v += __shfl_xor_sync(0xffffffffu, v, o); }
        if (g == 0) {
          const int ch = (a % NTF) * 8 + 2 * tq + e;
          atomicAdd(&stat_s[2 * ch], u);
          atomicAdd(&stat_s[2 * ch + 1], v);
        }
        s1[a][e] = 0.f; s2[a][e] = 0.f;
      }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * p.Cf; c += blockDim.x) {
      atomicAdd(&p.stats[(size_t)n_img * p.Cf * 2 + c], stat_s[c]);
      stat_s[c] = 0.f;
    }
    __syncthreads();
  };

  for (int it = 0; it < ntile; ++it) {
    const int s = it % p.stages;
    const int t = t_begin + it;
    const int tj = t % p.tiles_j, rr = t / p.tiles_j;
    const int ti = rr % p.tiles_i, n = rr / p.tiles_i;
    if (has_stats && n != stat_n) {
      if (stat_n >= 0) flush_stats(stat_n);
      stat_n = n;
    }
    mbar_wait(&full_bar[s], (uint32_t)((it / p.stages) & 1));
    const uint32_t x_base = smem_u32(smem + (size_t)s * p.x_bytes);

    for (int grp = warp; grp * MB < mt_tile; grp += CS_WARPS) {
      float acc[MB][NT][4];
#pragma unroll
      for (int a = 0; a < MB; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
      int il[MB], j0[MB];
#pragma unroll
      for (int a = 0; a < MB; ++a) {
        const int mt = grp * MB + a;
        il[a] = mt / mt_j; j0[a] = (mt - il[a] * mt_j) << 4;
      }
#pragma unroll
      for (int f = 0; f < F; ++f) {
        for (int tt = 0; tt < p.tpf; ++tt) {
          const int tap = f * p.tpf + tt;
          const int dh = p.tap_dh[tap] - p.dh_min, dw = p.tap_dw[tap] - p.dw_min;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t bf[NTF / 2][4];
#pragma unroll
            for (int q = 0; q < NTF / 2; ++q)
              cs_ldsm_x4(w_base + cs_swz<PX>((uint32_t)(tap * p.Cf + q * 16 + b_row), (uint32_t)(2 * ks + b_kc)), bf[q]);
#pragma unroll
            for (int a = 0; a < MB; ++a) {
              uint32_t af[4];
              const uint32_t pix = (uint32_t)((il[a] * p.sh + dh) * p.xbox_w + (j0[a] + a_row) * p.sw + dw);
              cs_ldsm_x4(x_base + cs_swz<PX>(pix, (uint32_t)(2 * ks + a_kc)), af);
#pragma unroll
              for (int q = 0; q < NTF; ++q)
                cs_mma(acc[a][f * NTF + q], af, bf[q >> 1][(q & 1) * 2], bf[q >> 1][(q & 1) * 2 + 1]);
            }
          }
        }
      }
      // ---- epilogue on the accumulator fragments: rows g / g+8 of each m-tile, channels 8*nt + 2*tq + {0,1}
#pragma unroll
      for (int a = 0; a < MB; ++a) {
        const int ho = ti * p.TI + il[a];
        const int wo0 = tj * p.TJ + j0[a] + g;
        const bool rowok = ho < p.Ho;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int wo = wo0 + 8 * h;
          const bool valid = rowok && wo < p.Wo;
          __nv_bfloat16* ypix = p.y + ((long long)n * p.ysn + (long long)ho * p.ysh + (long long)wo * p.ysw) + 2 * tq;
          // pair index of (this pixel, channel 2tq) in the logical [N,Ho,Wo,Cf] output of a fold
          const unsigned long long pair0 =
              (((unsigned long long)n * p.Ho + ho) * p.Wo + wo) * (unsigned long long)half_cf + tq;
          const float* zpix = noise_mode == 1
              ? p.noise + ((long long)n * p.zsn + (long long)ho * p.zsh + (long long)wo * p.zsw) + 2 * tq : nullptr;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            constexpr int dummy = 0; (void)dummy;
            const int f = nt / NTF, q = nt % NTF;      // compile-time after unrolling
            const int c = f * p.Cf + q * 8 + 2 * tq;   // channel in the launch (bias / noise weight index)
            const float2 bb = *reinterpret_cast<const float2*>(bias_s + c);
            float v0 = acc[a][nt][2 * h] + bb.x, v1 = acc[a][nt][2 * h + 1] + bb.y;
            if (noise_mode == 2) {
              const float2 z = normal_pair(nkey, pair0 + 4 * q);
              const float2 ww = *reinterpret_cast<const float2*>(nw_s + c);
              v0 = fmaf(ww.x, z.x, v0); v1 = fmaf(ww.y, z.y, v1);
            } else if (noise_mode == 1) {
              if (valid) {
                v0 = fmaf(nw_s[c], zpix[f * p.Cf + q * 8], v0); v1 = fmaf(nw_s[c + 1], zpix[f * p.Cf + q * 8 + 1], v1);
              }
            }
            if (act == HWG_ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            else if (act == HWG_ACT_LRELU) { v0 = fmaxf(v0, v0 * p.slope); v1 = fmaxf(v1, v1 * p.slope); }
            if (valid) {
              *reinterpret_cast<__nv_bfloat162*>(ypix + fold_off[f] + q * 8) = __floats2bfloat162_rn(v0, v1);
              if constexpr (STATS != 0) {
                s1[nt][0] += v0; s1[nt][1] += v1;
                s2[nt][0] = fmaf(v0, v0, s2[nt][0]); s2[nt][1] = fmaf(v1, v1, s2[nt][1]);
              }
            }
          }
        }
      }
    }
    __syncthreads();   // every warp is done reading stage s
    if (threadIdx.x == 0 && it + p.stages < ntile) issue(it + p.stages);
  }
  if (has_stats && stat_n >= 0) flush_stats(stat_n);
}

typedef CUresult (*PFN_encodeTiledC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);
static PFN_encodeTiledC cs_get_encode() {
  static PFN_encodeTiledC fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledC>(f);
  });
  return fn;
}
static int cs_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

typedef void (*CsKernel)(const CUtensorMap, const CUtensorMap, const CsParams);

// Launches the staged-tile kernel when the layer is one it covers; returns -1 otherwise (the caller then uses the
// tcgen05 kernel).
int conv_small_try(const hwgConvDesc* d, const void* x, const void* w, const float* bias, const float* noise,
                   const float* noise_w, float* stats, void* y, void* stream) {
  if (d->y_dtype != HWG_DT_BF16 || d->act == HWG_ACT_LOGSOFTMAX) return -1;
  static const bool env_off = getenv("HWG_NO_SMALL_CONV") != nullptr;   // development A/B switch
  if (env_off && !(d->fold_c && d->fold_taps > 0)) return -1;
  if (!(d->Cin == 16 || d->Cin == 32 || d->Cin == 64)) return -1;
  const int F = (d->fold_c && d->fold_taps > 0) ? d->Cout / d->fold_c : 1;
  if (d->fold_c && d->fold_taps <= 0) return -1;          // union-tap folding belongs to the tcgen05 kernel
  const int Cf = d->fold_c ? d->fold_c : d->Cout;
  if (Cf % 16 != 0) return -1;
  const int NTF = Cf / 8;
  CsKernel k = nullptr; int MB = 0;
  const int cin = d->Cin;
  const bool st = stats != nullptr;
  // the two epilogues the train step uses get their own instantiations: generator forward (in-kernel noise + LeakyReLU +
  // statistics) and plain (input gradients, up-sampling convolutions)
  const int nmode = noise_w ? (noise ? 1 : 2) : 0;
  const bool gen_fwd = st && nmode == 2 && d->act == HWG_ACT_LRELU;
  const bool plain = !st && nmode == 0 && d->act == HWG_ACT_NONE;
#define CS_PICK(CI, NF, FF, MBS, MBN)                                                              \
  if (cin == CI && NTF == NF && F == FF) {                                                         \
    if (st) { k = gen_fwd ? conv_small_kernel<CI, NF, FF, MBS, 1, HWG_ACT_LRELU, 2>                \
                          : conv_small_kernel<CI, NF, FF, MBS, 1>; MB = MBS; }                     \
    else { k = plain ? conv_small_kernel<CI, NF, FF, MBN, 0, HWG_ACT_NONE, 0>                      \
                     : conv_small_kernel<CI, NF, FF, MBN, 0>; MB = MBN; }                          \
  }
  CS_PICK(16, 2, 1, 4, 4)    // 16 -> 16  (block 4 conv2 and its input gradient)
  CS_PICK(16, 4, 1, 2, 4)    // 16 -> 32  (input gradient of block 4's FusedUpsample, stride 2)
  CS_PICK(32, 4, 1, 2, 4)    // 32 -> 32  (block 3 conv2 and its input gradient)
  CS_PICK(32, 2, 1, 4, 4)    // 32 -> 16
  CS_PICK(32, 8, 1, 1, 2)    // 32 -> 64  (input gradient of block 3's FusedUpsample, stride 2)
  CS_PICK(32, 2, 4, 1, 2)    // 32 -> 4 x 16 (block 4's FusedUpsample, parities as folds)
  CS_PICK(64, 4, 4, 1, 1)    // 64 -> 4 x 32 (block 3's FusedUpsample)
#undef CS_PICK
  if (!k) return -1;
  if (F > 1 && st) return -1;   // folded launches carry no statistics (the blur pass that follows computes them)
  if (F > 1 && (noise_w != nullptr || d->ntaps != F * d->fold_taps)) return -1;
  if ((reinterpret_cast<uintptr_t>(y) & 3) || (d->y_stride_w & 1) || (d->y_stride_h & 1) || (d->y_stride_n & 1) ||
      (d->fold_stride_h & 1) || (d->fold_stride_w & 1)) return -1;
  const int sh = d->in_stride_h > 1 ? d->in_stride_h : 1, sw = d->in_stride_w > 1 ? d->in_stride_w : 1;
  if (sh > 2 || sw > 2) return -1;
  int dh_min = 1 << 30, dh_max = -(1 << 30), dw_min = 1 << 30, dw_max = -(1 << 30);
  for (int t = 0; t < d->ntaps; ++t) {
    dh_min = d->tap_dh[t] < dh_min ? d->tap_dh[t] : dh_min; dh_max = d->tap_dh[t] > dh_max ? d->tap_dh[t] : dh_max;
    dw_min = d->tap_dw[t] < dw_min ? d->tap_dw[t] : dw_min; dw_max = d->tap_dw[t] > dw_max ? d->tap_dw[t] : dw_max;
  }
  if (dh_max - dh_min > 4 || dw_max - dw_min > 8) return -1;
  PFN_encodeTiledC encode = cs_get_encode();
  if (!encode) { set_error("hwg_conv_fprop: cuTensorMapEncodeTiled unavailable"); return HWG_ERR_CUDA; }

  CsParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout; p.Cf = Cf;
  p.sh = sh; p.sw = sw; p.dh_min = dh_min; p.dw_min = dw_min;
  p.ntaps = d->ntaps; p.tpf = F > 1 ? d->fold_taps : d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) { p.tap_dh[t] = d->tap_dh[t]; p.tap_dw[t] = d->tap_dw[t]; }
  const int px = cin * 2;
  p.w_box_bytes = 256 * px;
  const int w_rows = d->ntaps * Cf;
  p.w_boxes = (w_rows + 255) / 256;
  const int w_box_rows = w_rows < 256 ? w_rows : 256;
  if (w_rows < 256) p.w_box_bytes = w_rows * px;
  p.w_bytes = (p.w_boxes * p.w_box_bytes + 1023) / 1024 * 1024;
  // tile: TJ x TI output pixels; TI*TJ/16 m-tiles must be a multiple of MB
  p.TJ = d->Wo >= 64 ? 64 : (d->Wo > 16 ? 32 : 16);
  // two CTAs per SM unless the resident weights are large (the 64-channel FusedUpsample: 64 KiB)
  const int ctas_per_sm = p.w_bytes > 40 * 1024 ? 1 : 2;
  const int budget = (ctas_per_sm == 2 ? 110 : 210) * 1024 - p.w_bytes - 4096;
  int TI = 8;
  auto xbytes = [&](int ti) { return ((ti - 1) * sh + 1 + dh_max - dh_min) * ((p.TJ - 1) * sw + 1 + dw_max - dw_min) * px; };
  while (TI > 1 && (TI / 2 >= d->Ho || ((xbytes(TI) + 1023) / 1024 * 1024) * 2 > budget)) TI >>= 1;
  while ((TI * (p.TJ / 16)) % MB != 0) TI <<= 1;
  p.TI = TI;
  p.xbox_w = (p.TJ - 1) * sw + 1 + dw_max - dw_min;
  const int xbox_h = (TI - 1) * sh + 1 + dh_max - dh_min;
  if (p.xbox_w > 256 || xbox_h > 256) return -1;
  p.tx_bytes = xbytes(TI);
  p.x_bytes = (p.tx_bytes + 1023) / 1024 * 1024;
  p.stages = budget / p.x_bytes;
  if (p.stages > 6) p.stages = 6;
  if (p.stages < 2) return -1;
  p.tiles_j = (d->Wo + p.TJ - 1) / p.TJ;
  p.tiles_i = (d->Ho + p.TI - 1) / p.TI;
  p.total_tiles = p.tiles_i * p.tiles_j * d->N;
  p.ysn = d->y_stride_n; p.ysh = d->y_stride_h; p.ysw = d->y_stride_w;
  p.fold_sh = d->fold_stride_h; p.fold_sw = d->fold_stride_w; p.fold_w = d->fold_w > 0 ? d->fold_w : 1;
  p.zsn = d->nz_stride_n; p.zsh = d->nz_stride_h; p.zsw = d->nz_stride_w;
  p.act = d->act; p.slope = d->slope;
  p.noise_mode = noise_w ? (noise ? 1 : 2) : 0;
  p.has_stats = stats != nullptr;
  p.bias = bias; p.noise = noise; p.noise_w = noise_w; p.stats = stats;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.noise_seed = d->noise_seed; p.noise_subseq = d->noise_subseq;
  p.noise_seed_dev = reinterpret_cast<const unsigned long long*>(d->noise_seed_dev);

  const CUtensorMapSwizzle swz = cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (cin == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmx, tmw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_pitch * 2, (cuuint64_t)d->W * d->x_pitch * 2,
                             (cuuint64_t)d->H * d->W * d->x_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)cin, (cuuint32_t)p.xbox_w, (cuuint32_t)xbox_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("hwg_conv_fprop(small): cuTensorMapEncodeTiled(x) failed (%d)", (int)r); return HWG_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)cin, (cuuint32_t)w_box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("hwg_conv_fprop(small): cuTensorMapEncodeTiled(w) failed (%d)", (int)r); return HWG_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.x_bytes + p.w_bytes + (size_t)(2 * F * NTF * 8 + 2 * NTF * 8) * sizeof(float) +
                      10 * sizeof(uint64_t) + 1024;
  HWG_SMEM_OPTIN(k);
  int grid = ctas_per_sm * cs_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  k<<<grid, CS_WARPS * 32, smem, (cudaStream_t)stream>>>(tmx, tmw, p);
  return check_launch("conv_small_kernel");
}

}  // namespace hwg
