// Weight gradient of the small-channel convolutions (Cout, Cin in {16, 32}) — HBM-bound (sm_100a).
//
//   dw[t][co][ci] = sum_{n,i,j} gy[n, i*sh + goh + ph_t, j*sw + gow + pw_t, co] * x[n, i + dh_t, j + dw_t, ci]
//
// These layers (generator blocks 3-4: 32 and 16 channels on 32x512 / 64x1024 images) contract over ~10^6 pixels
// into a few thousand outputs: arithmetic intensity ~70 FLOP/B, far below the tensor ridge, so the design goal
// is "read gy and x once at HBM speed".  The tcgen05 split-K kernel of hwg_wgrad.cu is the wrong tool here: its
// 128-row MMA tile is >= 75 % padding and every tap re-reads both operands (it ran at 1.2 ms for the 16-channel
// layer).  This kernel instead
//   * walks spatial tiles with persistent CTAs; per tile ONE TMA box of gy and ONE halo box of x land in shared
//     memory (swizzled, zero-filled outside the image = the forward's zero padding), multi-stage mbarrier ring;
//   * every tap of the convolution reads the same staged tiles at shifted pixel offsets: warps issue
//     ldmatrix.trans + mma.sync.m16n8k16 (pixels are the K dimension, both operands pixel-major) with all
//     (tap, co, ci) accumulators of a warp resident in registers across the CTA's whole tile range;
//   * one shared-memory reduction over the CTA's warps and one red.global.add per output element per CTA.
// All phases of an up-sampling convolution (FusedUpsample: stride 2, 16 taps) run in one launch: each tap carries
// its own gy phase, so gy is read once instead of once per parity.
#include "common.cuh"
#include "sm100.cuh"
#include <cuda.h>
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace hwg {
using namespace sm100;

struct WgSmallParams {
  int Cout, Cin, ntaps;
  int Hi, Wi;                  // iteration grid
  int TI, TJ;                  // tile of the iteration grid (TJ multiple of 16)
  int tiles_i, tiles_j, total_tiles;
  int gsh, gsw, goh, gow;      // gy stride / offset
  int dh_min, dw_min;          // halo origin of the x box
  int gbox_w, xbox_w;          // box widths in pixels
  int g_bytes, stage_bytes, stages;   // g_bytes: gy box rounded up to 1 KiB (= offset of the x box in a stage)
  int tx_bytes;                       // exact bytes one stage's two TMA boxes deliver
  int TG, KS;                  // tap groups x K slices = warps
  int tap_dh[HWG_MAX_TAPS], tap_dw[HWG_MAX_TAPS], tap_ph[HWG_MAX_TAPS], tap_pw[HWG_MAX_TAPS];
  float* dw;
};

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of (pixel, 16-byte chunk) inside a TMA box whose pixel rows are PITCH = 32/64 bytes, with the
// matching 32B/64B swizzle: address bits [4,4+log2(PITCH/16)) ^= bits [7,...)
template <int PITCH>
__device__ __forceinline__ uint32_t swz(uint32_t pix, uint32_t chunk) {
  const uint32_t o = pix * PITCH + chunk * 16u;
  return o ^ (((o >> 7) & (PITCH / 16u - 1u)) << 4);
}

// MT = Cout/16, NT = Cin/8, TPW = taps per warp, MINB = resident CTAs per SM the register budget is cut for.
// Occupancy decides here (ncu, round 2: the <1,2,9> launch held 8 warps per SM — 12 % — with 44 % of the stalls on
// fixed-latency dependencies and 23 % on ldmatrix results): the MINB = 2 variants give a warp fewer taps (24-32
// accumulator registers instead of 72-96), so that two CTAs of 12-16 warps fit an SM.
template <int MT, int NT, int TPW, int MINB>
__global__ void __launch_bounds__(512, MINB)
wgrad_small_kernel(const __grid_constant__ CUtensorMap tmap_gy, const __grid_constant__ CUtensorMap tmap_x,
                   const __grid_constant__ WgSmallParams p) {
  constexpr int PA = MT * 32, PB = NT * 16;  // bytes per pixel of gy / x in shared memory
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  float* red = reinterpret_cast<float*>(full_bar + 8);   // [ntaps][Cout][Cin]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nred = p.ntaps * p.Cout * p.Cin;

  // contiguous tile range of this CTA
  const int per = (p.total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(p.total_tiles, t_begin + per);
  const int ntile = t_end - t_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_gy);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < nred; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  auto issue = [&](int it) {
    const int t = t_begin + it, s = it % p.stages;
    const int tj = t % p.tiles_j, r = t / p.tiles_j;
    const int ti = r % p.tiles_i, n = r / p.tiles_i;
    unsigned char* dst = smem + (size_t)s * p.stage_bytes;
    mbar_expect_tx(&full_bar[s], (uint32_t)p.tx_bytes);
    tma_load_4d(dst, &tmap_gy, &full_bar[s], 0, tj * p.TJ * p.gsw + p.gow, ti * p.TI * p.gsh + p.goh, n);
    tma_load_4d(dst + p.g_bytes, &tmap_x, &full_bar[s], 0, tj * p.TJ + p.dw_min, ti * p.TI + p.dh_min, n);
  };
  if (threadIdx.x == 0)
    for (int it = 0; it < min(p.stages, ntile); ++it) issue(it);

  const int tg = warp % p.TG, ks = warp / p.TG;
  const int tap0 = tg * TPW;
  float acc[TPW][MT][NT][4];
#pragma unroll
  for (int a = 0; a < TPW; ++a)
#pragma unroll
    for (int b = 0; b < MT; ++b)
#pragma unroll
      for (int c = 0; c < NT; ++c)
#pragma unroll
        for (int d = 0; d < 4; ++d) acc[a][b][c][d] = 0.f;

  // per-lane ldmatrix roles
  const int mi = lane >> 3, rr = lane & 7;
  const int a_k = rr + 8 * (mi >> 1), a_chunk = mi & 1;   // A: matrices (k0-7,c0) (k0-7,c1) (k8-15,c0) (k8-15,c1)
  const int b_k = rr + 8 * (mi & 1), b_chunk = mi >> 1;   // B: matrices (k0-7,c0) (k8-15,c0) (k0-7,c1) (k8-15,c1)
  const int ksteps_j = p.TJ >> 4, ksteps = p.TI * ksteps_j;
  const int kj_shift = ksteps_j == 4 ? 2 : (ksteps_j == 2 ? 1 : 0);   // TJ is 16, 32 or 64

  // Per-warp tap constants, in registers: the round-2 profile of this loop showed ~140 instructions per K step for six
  // MMAs — an integer division per step, four dynamically indexed constant loads and a full address rebuild per tap.
  //   xoff: the tap's pixel offset inside the x halo box;  goff: its phase's pixel offset inside the gy box;
  //   reload bit tt: tap tt reads gy at another phase than tap tt-1 (bit 0 always set)
  int xoff[TPW], goff[TPW];
  unsigned reload = 0u;
  const int ntw = max(0, min(TPW, p.ntaps - tap0));
#pragma unroll
  for (int tt = 0; tt < TPW; ++tt) {
    const int tap = min(tap0 + tt, p.ntaps - 1);
    xoff[tt] = (p.tap_dh[tap] - p.dh_min) * p.xbox_w + (p.tap_dw[tap] - p.dw_min);
    goff[tt] = p.tap_ph[tap] * p.gbox_w + p.tap_pw[tap];
    if (tt == 0 || goff[tt] != goff[tt > 0 ? tt - 1 : 0]) reload |= 1u << tt;
  }

  for (int it = 0; it < ntile; ++it) {
    const int s = it % p.stages;
    mbar_wait(&full_bar[s], (uint32_t)((it / p.stages) & 1));
    const int t = t_begin + it;
    const int ti = (t / p.tiles_j) % p.tiles_i;
    const int i_lim = min(p.TI, p.Hi - ti * p.TI);   // iteration rows that exist in this tile
    const uint32_t g_base = smem_u32(smem + (size_t)s * p.stage_bytes);
    const uint32_t x_base = g_base + (uint32_t)p.g_bytes;
    for (int kq = ks; kq < ksteps; kq += p.KS) {
      const int il = kq >> kj_shift, j0 = (kq & (ksteps_j - 1)) << 4;
      if (il >= i_lim) break;
      const uint32_t gpix0 = (uint32_t)((il * p.gsh) * p.gbox_w + (j0 + a_k) * p.gsw);
      const uint32_t xpix0 = (uint32_t)(il * p.xbox_w + j0 + b_k);
      uint32_t afr[MT][4];
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        if (tt < ntw) {
          if ((reload >> tt) & 1u) {
            const uint32_t pix = gpix0 + (uint32_t)goff[tt];
#pragma unroll
            for (int m = 0; m < MT; ++m) ldsm_x4_trans(g_base + swz<PA>(pix, (uint32_t)(2 * m + a_chunk)), afr[m]);
          }
          const uint32_t xpix = xpix0 + (uint32_t)xoff[tt];
#pragma unroll
          for (int n2 = 0; n2 < NT / 2; ++n2) {
            uint32_t bfr[4];
            ldsm_x4_trans(x_base + swz<PB>(xpix, (uint32_t)(2 * n2 + b_chunk)), bfr);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              mma_bf16_16816(acc[tt][m][2 * n2], afr[m], bfr[0], bfr[1]);
              mma_bf16_16816(acc[tt][m][2 * n2 + 1], afr[m], bfr[2], bfr[3]);
            }
          }
        }
      }
    }
    __syncthreads();   // every warp is done with stage s
    if (threadIdx.x == 0 && it + p.stages < ntile) issue(it + p.stages);
  }

  // fold the K slices of the CTA in shared memory, then one global reduction per output element
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int tt = 0; tt < TPW; ++tt) {
    const int tap = tap0 + tt;
    if (tap < p.ntaps) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          float* base = red + ((size_t)tap * p.Cout + m * 16) * p.Cin + n * 8 + 2 * tq;
          atomicAdd(base + (size_t)g * p.Cin, acc[tt][m][n][0]);
          atomicAdd(base + (size_t)g * p.Cin + 1, acc[tt][m][n][1]);
          atomicAdd(base + (size_t)(g + 8) * p.Cin, acc[tt][m][n][2]);
          atomicAdd(base + (size_t)(g + 8) * p.Cin + 1, acc[tt][m][n][3]);
        }
    }
  }
  __syncthreads();
  if (ntile > 0)
    for (int i = threadIdx.x; i < nred; i += blockDim.x) atomicAdd(&p.dw[i], red[i]);
}

typedef CUresult (*PFN_encodeTiledS)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);
static PFN_encodeTiledS wgs_get_encode() {
  static PFN_encodeTiledS fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledS>(f);
  });
  return fn;
}

static int wgs_encode(PFN_encodeTiledS encode, CUtensorMap* tm, const void* base, int C, int W, int H, int N, int pitch,
                      int box_w, int box_h, const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle swz = C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("hwg_conv_wgrad(small): cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r); return HWG_ERR_CUDA; }
  return HWG_OK;
}

static int wgs_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

typedef void (*WgsKernel)(const CUtensorMap, const CUtensorMap, const WgSmallParams);

// Returns HWG_OK after launching, or -1 when the shape is not one this kernel covers (caller falls through to the
// tcgen05 split-K kernel).
int wgrad_small_try(const hwgWgradDesc* d, const void* x, const void* gy, float* dw, void* stream) {
  if (!((d->Cout == 16 || d->Cout == 32) && (d->Cin == 16 || d->Cin == 32))) return -1;
  const int Hi = d->Hi > 0 ? d->Hi : d->Ho, Wi = d->Wi > 0 ? d->Wi : d->Wo;
  const int gsh = d->gy_stride_h > 1 ? d->gy_stride_h : 1, gsw = d->gy_stride_w > 1 ? d->gy_stride_w : 1;
  if (gsh > 2 || gsw > 2) return -1;
  WgSmallParams p;
  memset(&p, 0, sizeof(p));
  int dh_min = 1 << 30, dh_max = -(1 << 30), dw_min = 1 << 30, dw_max = -(1 << 30);
  for (int t = 0; t < d->ntaps; ++t) {
    p.tap_dh[t] = d->tap_dh[t]; p.tap_dw[t] = d->tap_dw[t];
    p.tap_ph[t] = d->tap_gy_h[t]; p.tap_pw[t] = d->tap_gy_w[t];
    if (p.tap_ph[t] < 0 || p.tap_ph[t] >= gsh || p.tap_pw[t] < 0 || p.tap_pw[t] >= gsw) return -1;
    dh_min = d->tap_dh[t] < dh_min ? d->tap_dh[t] : dh_min; dh_max = d->tap_dh[t] > dh_max ? d->tap_dh[t] : dh_max;
    dw_min = d->tap_dw[t] < dw_min ? d->tap_dw[t] : dw_min; dw_max = d->tap_dw[t] > dw_max ? d->tap_dw[t] : dw_max;
  }
  if (dh_max - dh_min > 4 || dw_max - dw_min > 8) return -1;
  // columns past the iteration grid must fall outside gy (TMA zero fill); rows are bounded in the kernel
  if ((long)Wi * gsw + d->gy_off_w < d->Wo) return -1;
  PFN_encodeTiledS encode = wgs_get_encode();
  if (!encode) { set_error("hwg_conv_wgrad: cuTensorMapEncodeTiled unavailable"); return HWG_ERR_CUDA; }

  p.Cout = d->Cout; p.Cin = d->Cin; p.ntaps = d->ntaps;
  p.Hi = Hi; p.Wi = Wi; p.gsh = gsh; p.gsw = gsw; p.goh = d->gy_off_h; p.gow = d->gy_off_w;
  p.dh_min = dh_min; p.dw_min = dw_min;
  const int MT = d->Cout / 16, NT = d->Cin / 8;
  static const int variant_env = getenv("HWG_WGS_VARIANT") ? atoi(getenv("HWG_WGS_VARIANT")) : 1;   // development A/B switch
  const int pa = d->Cout * 2, pb = d->Cin * 2;
  const int nred_bytes = d->ntaps * d->Cout * d->Cin * 4;
  WgsKernel k = nullptr;
  int ctas_per_sm = 1;
  // variant 1: few taps per warp, two CTAs per SM; variant 0: all accumulators of a tap group in one warp, one CTA
  // per SM (also the fall-back when two CTAs' stages do not fit next to the reduction tiles)
  auto configure = [&](int variant) -> bool {
    int TPW;
    if (variant == 0) {
      ctas_per_sm = 1;
      if (MT == 1 && NT == 2) { TPW = 9; k = wgrad_small_kernel<1, 2, 9, 1>; }
      else if (MT == 1 && NT == 4) { TPW = 4; k = wgrad_small_kernel<1, 4, 4, 1>; }
      else if (MT == 2 && NT == 2) { TPW = 4; k = wgrad_small_kernel<2, 2, 4, 1>; }
      else { TPW = 3; k = wgrad_small_kernel<2, 4, 3, 1>; }
    } else {
      ctas_per_sm = 2;
      if (MT == 1 && NT == 2) { TPW = 3; k = wgrad_small_kernel<1, 2, 3, 2>; }
      else if (MT == 1 && NT == 4) { TPW = 2; k = wgrad_small_kernel<1, 4, 2, 2>; }
      else if (MT == 2 && NT == 2) { TPW = 2; k = wgrad_small_kernel<2, 2, 2, 2>; }
      else { TPW = 1; k = wgrad_small_kernel<2, 4, 1, 2>; }
    }
    p.TG = (d->ntaps + TPW - 1) / TPW;
    if (p.TG > 16) return false;
    p.KS = 16 / p.TG;
    if (variant == 0 && p.KS > 8) p.KS = 8;
    if (p.KS < 1) p.KS = 1;
    p.TJ = Wi >= 64 ? 64 : (Wi > 16 ? 32 : 16);
    auto stage_bytes_for = [&](int ti, int* gb) {
      const int g = ((ti * gsh) * (p.TJ * gsw) * pa + 1023) / 1024 * 1024;
      const int xb = ((ti + dh_max - dh_min) * (p.TJ + dw_max - dw_min) * pb + 1023) / 1024 * 1024;
      if (gb) *gb = g;
      return g + xb;
    };
    // two resident CTAs share the SM's shared memory: two stages of each must fit next to the reduction tile
    const int per_cta = ctas_per_sm == 2 ? (220 * 1024) / 2 - nred_bytes - 2048 : 200 * 1024 - nred_bytes - 2048;
    int stage_limit = 48 * 1024;   // a stage near 40-60 KB
    if (ctas_per_sm == 2 && per_cta / 2 < stage_limit) stage_limit = per_cta / 2;
    int TI = 8;
    while (TI > 1 && (stage_bytes_for(TI, nullptr) > stage_limit || TI / 2 >= Hi)) TI >>= 1;
    p.TI = TI;
    p.stage_bytes = stage_bytes_for(TI, &p.g_bytes);
    const int max_stages = ctas_per_sm == 2 ? 3 : 4;
    p.stages = per_cta / p.stage_bytes;
    if (p.stages > max_stages) p.stages = max_stages;
    return p.stages >= 2;
  };
  if (!((variant_env != 0 && configure(1)) || configure(0))) return -1;
  p.gbox_w = p.TJ * gsw; p.xbox_w = p.TJ + dw_max - dw_min;
  p.tiles_j = (Wi + p.TJ - 1) / p.TJ;
  p.tiles_i = (Hi + p.TI - 1) / p.TI;
  p.total_tiles = p.tiles_i * p.tiles_j * d->N;
  // expect_tx counts the full boxes (zero-filled parts included)
  p.tx_bytes = (p.TI + dh_max - dh_min) * p.xbox_w * pb + (p.TI * gsh) * p.gbox_w * pa;
  p.dw = dw;

  CUtensorMap tmg, tmx;
  int rc = wgs_encode(encode, &tmg, gy, d->Cout, d->Wo, d->Ho, d->N, d->gy_pitch, p.gbox_w, p.TI * gsh, "gy");
  if (rc) return rc;
  rc = wgs_encode(encode, &tmx, x, d->Cin, d->W, d->H, d->N, d->x_pitch, p.xbox_w, p.TI + dh_max - dh_min, "x");
  if (rc) return rc;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 64 + nred_bytes + 1024;
  HWG_SMEM_OPTIN(k);
  int grid = wgs_sms() * ctas_per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  k<<<grid, 32 * p.TG * p.KS, smem, (cudaStream_t)stream>>>(tmg, tmx, p);
  return check_launch("wgrad_small_kernel");
}

}  // namespace hwg
