// Convolution weight gradient on tcgen05 (sm_100a).
//
//   dw[t][co][ci] = sum_pixels gy[pix][co] * x[pix + off_t][ci]
//
//   GEMM view: D[M = 128 output channels][N = up to 256 input channels] += A * B with
//   K = pixels.  Both operands are "MN-major": a TMA box {64 channels, PW x PH pixels} of an NHWC
//   tensor lands as rows of 128 bytes (64 channels) per pixel with the 128-byte swizzle, which is
//   the canonical MN-major SWIZZLE_128B UMMA layout (8-pixel groups 1 KiB apart = SBO, 64-channel
//   blocks one box apart = LBO).  The x box is read at the tap's pixel offset; out-of-bounds pixels
//   arrive as zeros (= the forward's zero padding), out-of-range gy pixels as zeros too.
//
//   One CTA = one (tap, co tile, ci tile) and one contiguous range of 64-pixel chunks (split-K);
//   warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, warps 2-5 = epilogue that adds the
//   fp32 accumulator into dw with red.global.add.v4.f32.
#include "common.cuh"
#include <stdlib.h>
#include "sm100.cuh"
#include <cuda.h>
#include <mutex>
#include <string.h>

namespace hwg {
using namespace sm100;

constexpr int WG_PIX = 64;  // pixels per K chunk

struct WgradKParams {
  int Cout, Cin, ntaps;
  int PW, PH, chunks_w, chunks_h, total_chunks, chunks_per_split;
  int co_tiles, ci_tiles, BN;  // BN = ci tile width (<= 256)
  int CBa, CBb;                // channels per TMA box of gy / x (64, 32 or 16 -> 128/64/32-byte swizzle)
  int gsh, gsw, goh, gow;      // where iteration-grid point (i,j) sits in gy: (i*gsh+goh, j*gsw+gow)
  int stages, a_bytes, b_bytes, tmem_cols;
  int tap_dh[HWG_MAX_TAPS], tap_dw[HWG_MAX_TAPS];
  float* dw;
};

// MN-major operand made of TMA boxes {CB channels, 64 pixels}: a pixel row is CB*2 = 128/64/32 bytes (the
// swizzle span), 8-pixel groups are 8 rows apart (SBO), CB-channel blocks one box apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr, uint32_t row_bytes, uint32_t lbo_bytes) {
  const uint32_t mode = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)mode << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(192)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_gy, const __grid_constant__ CUtensorMap tmap_x,
                  const __grid_constant__ WgradKParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  int unit = blockIdx.y;
  const int ci_t = unit % p.ci_tiles; unit /= p.ci_tiles;
  const int co_t = unit % p.co_tiles; unit /= p.co_tiles;
  const int tap = unit;
  const int co0 = co_t * 128, ci0 = ci_t * p.BN;
  const int c_begin = blockIdx.x * p.chunks_per_split;
  const int c_end = min(p.total_chunks, c_begin + p.chunks_per_split);
  const int kiters = c_end - c_begin;
  const int nblk_a = 128 / p.CBa, nblk_b = p.BN / p.CBb;
  const uint32_t boxa_bytes = WG_PIX * p.CBa * 2, boxb_bytes = WG_PIX * p.CBb * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_gy);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (kiters <= 0) {
    // nothing to do for this split (grid rounded up); still release TMEM below
  } else if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
      for (int c = c_begin; c < c_end; ++c) {
        const int wc = c % p.chunks_w, r = c / p.chunks_w;
        const int hc = r % p.chunks_h, n = r / p.chunks_h;
        const int w0 = wc * p.PW, h0 = hc * p.PH;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        unsigned char* a_dst = smem + (size_t)stage * stage_bytes;
        unsigned char* b_dst = a_dst + p.a_bytes;
        mbar_expect_tx(&full_bar[stage], (uint32_t)(nblk_a * boxa_bytes + nblk_b * boxb_bytes));
        for (int b = 0; b < nblk_a; ++b)   // channel blocks past Cout arrive as zeros
          tma_load_4d(a_dst + b * boxa_bytes, &tmap_gy, &full_bar[stage], co0 + p.CBa * b, w0 * p.gsw + p.gow,
                      h0 * p.gsh + p.goh, n);
        for (int b = 0; b < nblk_b; ++b)
          tma_load_4d(b_dst + b * boxb_bytes, &tmap_x, &full_bar[stage], ci0 + p.CBb * b, w0 + dw, h0 + dh, n);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // both operands MN-major: a_major (bit 15) = b_major (bit 16) = 1
      const uint32_t idesc = umma_idesc_bf16(128, p.BN) | (1u << 15) | (1u << 16);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t b_addr = a_addr + (uint32_t)p.a_bytes;
        const uint64_t da = umma_desc_mnmajor(a_addr, (uint32_t)p.CBa * 2u, boxa_bytes);
        const uint64_t db = umma_desc_mnmajor(b_addr, (uint32_t)p.CBb * 2u, boxb_bytes);
        // 16 pixels = two 8-row groups = 16 rows further into the box (start address is in 16-byte units)
        const uint64_t ka = (uint64_t)(p.CBa * 2), kb = (uint64_t)(p.CBb * 2);
#pragma unroll
        for (int kk = 0; kk < WG_PIX / 16; ++kk)
          umma_bf16(tmem_base, da + kk * ka, db + kk * kb, idesc, (it | kk) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (it == kiters - 1) umma_commit(tmem_full);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float* drow = p.dw + ((size_t)tap * p.Cout + co) * p.Cin + ci0;
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(trow + (uint32_t)c0, r);
      tmem_ld_wait();
      if (co < p.Cout) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (c0 + j < p.BN)
            red_add_v4(drow + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                       __uint_as_float(r[j + 3]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static PFN_encodeTiled wg_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

static int encode_nhwc(PFN_encodeTiled encode, CUtensorMap* tm, const void* base, int C, int W, int H, int N,
                       int pitch, int CB, int PW, int PH, int sw, int sh, const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)(PW * sw), (cuuint32_t)(PH * sh), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
  const CUtensorMapSwizzle swz = CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (CB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("hwg_conv_wgrad: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r); return HWG_ERR_CUDA; }
  return HWG_OK;
}

int wgrad_small_try(const hwgWgradDesc* d, const void* x, const void* gy, float* dw, void* stream);  // hwg_wgrad_small.cu

}  // namespace hwg

using namespace hwg;

extern "C" int hwg_conv_wgrad(const hwgWgradDesc* d, const void* x, const void* gy, float* dw, void* stream) {
  HWG_REQUIRE(d && x && gy && dw, "hwg_conv_wgrad: null pointer");
  HWG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0, "hwg_conv_wgrad: empty extent");
  HWG_REQUIRE(d->Cin == 16 || d->Cin == 32 || (d->Cin >= 64 && d->Cin % 64 == 0),
              "hwg_conv_wgrad: Cin=%d must be 16, 32 or a multiple of 64", d->Cin);
  HWG_REQUIRE(d->Cout == 16 || d->Cout == 32 || (d->Cout >= 64 && d->Cout % 8 == 0),
              "hwg_conv_wgrad: Cout=%d must be 16, 32 or a multiple of 8 >= 64", d->Cout);
  HWG_REQUIRE(d->x_pitch >= d->Cin && d->x_pitch % 8 == 0 && d->gy_pitch >= d->Cout && d->gy_pitch % 8 == 0,
              "hwg_conv_wgrad: bad channel pitch");
  HWG_REQUIRE(d->ntaps >= 1 && d->ntaps <= HWG_MAX_TAPS, "hwg_conv_wgrad: ntaps=%d", d->ntaps);
  HWG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(dw) & 15) == 0, "hwg_conv_wgrad: pointers must be 16-byte aligned");
  // small-channel layers: the HBM-bound staged-tile kernel (hwg_wgrad_small.cu)
  {
    const int rc = wgrad_small_try(d, x, gy, dw, stream);
    if (rc >= 0) return rc;
  }
  for (int t = 0; t < d->ntaps; ++t)
    HWG_REQUIRE(d->tap_gy_h[t] == 0 && d->tap_gy_w[t] == 0,
                "hwg_conv_wgrad: per-tap gy phases are only supported for Cout, Cin in {16, 32}");
  PFN_encodeTiled encode = wg_get_encode();
  if (!encode) { set_error("hwg_conv_wgrad: cuTensorMapEncodeTiled unavailable"); return HWG_ERR_CUDA; }

  WgradKParams p;
  memset(&p, 0, sizeof(p));
  p.Cout = d->Cout; p.Cin = d->Cin; p.ntaps = d->ntaps;
  // 64-pixel K chunks: as wide as the output row allows
  const int Hi = d->Hi > 0 ? d->Hi : d->Ho, Wi = d->Wi > 0 ? d->Wi : d->Wo;
  p.gsh = d->gy_stride_h > 1 ? d->gy_stride_h : 1;
  p.gsw = d->gy_stride_w > 1 ? d->gy_stride_w : 1;
  p.goh = d->gy_off_h; p.gow = d->gy_off_w;
  HWG_REQUIRE(p.gsh <= 4 && p.gsw <= 4, "hwg_conv_wgrad: gy strides up to 4");
  int PW = 64;
  while (PW > 8 && PW / 2 >= Wi) PW >>= 1;
  p.PW = PW; p.PH = WG_PIX / PW;
  p.chunks_w = (Wi + p.PW - 1) / p.PW;
  p.chunks_h = (Hi + p.PH - 1) / p.PH;
  p.total_chunks = p.chunks_w * p.chunks_h * d->N;
  p.co_tiles = (d->Cout + 127) / 128;
  p.BN = d->Cin % 256 == 0 ? 256 : (d->Cin % 128 == 0 ? 128 : (d->Cin >= 64 ? 64 : d->Cin));
  p.ci_tiles = d->Cin / p.BN;
  p.CBa = d->Cout >= 64 ? 64 : d->Cout;
  p.CBb = d->Cin >= 64 ? 64 : d->Cin;
  p.a_bytes = 128 * WG_PIX * 2;            // always 128 output-channel rows (blocks past Cout are zero-filled)
  p.b_bytes = p.BN * WG_PIX * 2;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  p.stages = (190 * 1024) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  p.tmem_cols = p.BN <= 32 ? 32 : (p.BN <= 64 ? 64 : (p.BN <= 128 ? 128 : 256));
  for (int t = 0; t < d->ntaps; ++t) { p.tap_dh[t] = d->tap_dh[t]; p.tap_dw[t] = d->tap_dw[t]; }
  p.dw = dw;
  // split the pixel range so that the grid covers the SMs ~2x
  const int units = d->ntaps * p.co_tiles * p.ci_tiles;
  int cta_target = 2 * 148;
  if (const char* ov = getenv("HWG_WGRAD_CTAS")) {   // development override (tools/step_runner.py sweeps)
    const int v = atoi(ov);
    if (v > 0) cta_target = v;
  }
  int splits = (cta_target + units - 1) / units;
  if (splits > p.total_chunks) splits = p.total_chunks;
  if (splits < 1) splits = 1;
  p.chunks_per_split = (p.total_chunks + splits - 1) / splits;
  splits = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;

  CUtensorMap tmg, tmx;
  int rc = encode_nhwc(encode, &tmg, gy, d->Cout, d->Wo, d->Ho, d->N, d->gy_pitch, p.CBa, p.PW, p.PH, p.gsw, p.gsh, "gy");
  if (rc) return rc;
  rc = encode_nhwc(encode, &tmx, x, d->Cin, d->W, d->H, d->N, d->x_pitch, p.CBb, p.PW, p.PH, 1, 1, "x");
  if (rc) return rc;
  const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * sizeof(uint64_t) + 16 + 1024;
  HWG_SMEM_OPTIN(conv_wgrad_kernel);
  dim3 grid((unsigned)splits, (unsigned)units);
  conv_wgrad_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(tmg, tmx, p);
  return check_launch("conv_wgrad_kernel");
}
