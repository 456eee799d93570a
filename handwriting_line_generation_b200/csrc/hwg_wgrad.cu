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

struct WgradKParams {
  int Cout, Cin, ntaps;
  int PW, PH, pix;             // K chunk = PH x PW pixels of the iteration grid (pix = 64 or 128)
  int chunks_w, chunks_h, total_chunks, chunks_per_split;
  int co_tiles, ci_tiles, BN;  // BN = ci tile width (<= 256)
  int CBa, CBb;                // channels per TMA box of gy / x (64, 32 or 16 -> 128/64/32-byte swizzle)
  int nblk_a;                  // gy channel blocks that exist in the widest co tile (blocks past Cout are not loaded)
  int gsh, gsw, goh, gow;      // where iteration-grid point (i,j) sits in gy: (i*gsh+goh, j*gsw+gow)
  int tpc, tap_groups;         // taps per CTA (each owns BN accumulator columns), number of tap groups
  int halo;                    // 1: ONE x box {64 ch, PW+span_w, PH+span_h} per ci block, taps = shifted views of it
  int hb_w, hb_h, hb_bytes;    // halo box extent in pixels, bytes per ci block (rounded up to 1 KiB)
  int dh_min, dw_min;
  int stages, a_bytes, b_bytes, tmem_cols, unit_major;
  int tap_dh[HWG_MAX_TAPS], tap_dw[HWG_MAX_TAPS];
  float* dw;
};

// MN-major operands made of TMA boxes {CB channels, pixels}: a pixel row is CB*2 = 128/64/32 bytes (the swizzle span),
// 8-pixel groups are 8 rows apart (SBO), CB-channel blocks one box apart (LBO): sm100::umma_desc_lo / umma_desc_hi.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// One CTA = one (tap group, co tile, ci tile) and one contiguous range of pixel chunks (split-K).  The gy tile of a
// chunk is staged ONCE for all taps of the group; x arrives either as one box per tap or — halo mode — as one box
// that covers every tap's shifted window (the round-1 kernel staged both operands once per tap and was bound by the
// L2 -> shared-memory stream: 5.2 TB/s at 10-23 % tensor-pipe activity).  Tap t of the group accumulates into TMEM
// columns [t*BN, (t+1)*BN).
__global__ void __launch_bounds__(192)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_gy, const __grid_constant__ CUtensorMap tmap_x,
                  const __grid_constant__ WgradKParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  // barriers sit in front of the ring: the last stage's A descriptor may span rows past its loaded blocks (never
  // read back), which must still be inside the allocation
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  unsigned char* ring = smem + 1024;

  // unit-major launch order (p.unit_major): the (tap group, co tile, ci tile) units of one pixel range are
  // neighbours in the launch order, so they run at the same time and share their gy / x boxes through L2
  int unit = p.unit_major ? blockIdx.x : blockIdx.y;
  const int split_idx = p.unit_major ? blockIdx.y : blockIdx.x;
  const int ci_t = unit % p.ci_tiles; unit /= p.ci_tiles;
  const int co_t = unit % p.co_tiles; unit /= p.co_tiles;
  const int tap0 = unit * p.tpc;
  const int ntap = min(p.tpc, p.ntaps - tap0);
  const int co0 = co_t * 128, ci0 = ci_t * p.BN;
  const int c_begin = split_idx * p.chunks_per_split;
  const int c_end = min(p.total_chunks, c_begin + p.chunks_per_split);
  const int kiters = c_end - c_begin;
  const int nblk_a = min(p.nblk_a, (p.Cout - co0 + p.CBa - 1) / p.CBa), nblk_b = p.BN / p.CBb;
  const uint32_t boxa_bytes = p.pix * p.CBa * 2, boxb_bytes = p.pix * p.CBb * 2;

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
  if (kiters <= 0 || ntap <= 0) {
    // nothing to do for this split (grid rounded up); still release TMEM below
  } else if (warp == 0) {
    // TMA producer: the whole warp walks the loop, the elected lane issues (see sm100::elect_one)
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    const uint32_t tx = (uint32_t)nblk_a * boxa_bytes +
                        (p.halo ? (uint32_t)(nblk_b * p.hb_w * p.hb_h * p.CBb * 2)
                                : (uint32_t)(ntap * nblk_b) * boxb_bytes);
    for (int c = c_begin; c < c_end; ++c) {
      const int wc = c % p.chunks_w, r = c / p.chunks_w;
      const int hc = r % p.chunks_h, n = r / p.chunks_h;
      const int w0 = wc * p.PW, h0 = hc * p.PH;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (leader) {
        unsigned char* a_dst = ring + (size_t)stage * stage_bytes;
        unsigned char* b_dst = a_dst + p.a_bytes;
        mbar_expect_tx(&full_bar[stage], tx);
        for (int b = 0; b < nblk_a; ++b)
          tma_load_4d(a_dst + b * boxa_bytes, &tmap_gy, &full_bar[stage], co0 + p.CBa * b, w0 * p.gsw + p.gow,
                      h0 * p.gsh + p.goh, n);
        if (p.halo) {
          for (int b = 0; b < nblk_b; ++b)
            tma_load_4d(b_dst + (size_t)b * p.hb_bytes, &tmap_x, &full_bar[stage], ci0 + p.CBb * b, w0 + p.dw_min,
                        h0 + p.dh_min, n);
        } else {
          for (int t = 0; t < ntap; ++t)
            for (int b = 0; b < nblk_b; ++b)
              tma_load_4d(b_dst + (size_t)(t * nblk_b + b) * boxb_bytes, &tmap_x, &full_bar[stage], ci0 + p.CBb * b,
                          w0 + p.tap_dw[tap0 + t], h0 + p.tap_dh[tap0 + t], n);
        }
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp in the loop, elected lane issues; descriptors as (lo, hi) words, only the start-address
    // field of the low words moves (see hwg_conv.cu: the issuing lane is the serial resource of the small-N layers)
    const bool leader = elect_one();
    // both operands MN-major: a_major (bit 15) = b_major (bit 16) = 1
    const uint32_t idesc = umma_idesc_bf16(128, p.BN) | (1u << 15) | (1u << 16);
    const uint32_t rowa = (uint32_t)p.CBa * 2u, rowb = (uint32_t)p.CBb * 2u;
    const uint32_t hi_a = umma_desc_hi(8u * rowa, rowa), hi_b = umma_desc_hi(8u * rowb, rowb);
    const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFFu) >> 4;
    const uint32_t lbo_a = ((boxa_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t lbo_b = (((p.halo ? (uint32_t)p.hb_bytes : boxb_bytes) >> 4) & 0x3FFFu) << 16;
    const uint32_t stage16 = (uint32_t)stage_bytes >> 4, a16 = (uint32_t)p.a_bytes >> 4;
    const int segs = p.PW >> 4;           // halo mode: 16-pixel K steps per tile row
    const int ksteps = p.pix >> 4;
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < kiters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_lo = (ring_lo + (uint32_t)stage * stage16) | lbo_a;
      const uint32_t b_lo0 = (ring_lo + (uint32_t)stage * stage16 + a16) | lbo_b;
      if (p.halo) {
        for (int t = 0; t < ntap; ++t) {
          const int oh = p.tap_dh[tap0 + t] - p.dh_min, ow = p.tap_dw[tap0 + t] - p.dw_min;
          const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.BN);
          for (int i = 0; i < p.PH; ++i)
            for (int sgm = 0; sgm < segs; ++sgm) {
              // K step = pixels (i, 16*sgm .. 16*sgm+15): rows i*PW + 16*sgm of the gy tile, rows
              // (i+oh)*hb_w + 16*sgm + ow of the halo tile (the swizzle follows the absolute shared-memory
              // address, so a view may start at any 128-byte row: tools/halo_probe.cu)
              const uint32_t ra = (uint32_t)(i * p.PW + 16 * sgm);
              const uint32_t rb = (uint32_t)((i + oh) * p.hb_w + 16 * sgm + ow);
              if (leader)
                umma_bf16_lh(d_tmem, a_lo + ((ra * rowa) >> 4), hi_a, b_lo0 + ((rb * rowb) >> 4), hi_b, idesc,
                             (it | i | sgm) != 0 ? 1u : 0u);
            }
        }
      } else {
        const uint32_t tap16 = (uint32_t)nblk_b * (boxb_bytes >> 4);
        for (int t = 0; t < ntap; ++t) {
          const uint32_t b_lo = b_lo0 + (uint32_t)t * tap16;
          const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.BN);
          // 16 pixels = two 8-row groups = 16 rows further into the box (start address is in 16-byte units)
          if (leader) {
            if (ksteps == 4) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16_lh(d_tmem, a_lo + (uint32_t)kk * rowa, hi_a, b_lo + (uint32_t)kk * rowb, hi_b, idesc,
                             (it | kk) != 0 ? 1u : 0u);
            } else {
              for (int kk = 0; kk < ksteps; ++kk)
                umma_bf16_lh(d_tmem, a_lo + (uint32_t)kk * rowa, hi_a, b_lo + (uint32_t)kk * rowb, hi_b, idesc,
                             (it | kk) != 0 ? 1u : 0u);
            }
          }
        }
      }
      if (leader) {
        umma_commit(&empty_bar[stage]);
        if (it == kiters - 1) umma_commit(tmem_full);
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    if (co0 + q * 32 < p.Cout) {          // warp-uniform: quarters past Cout hold nothing
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int t = 0; t < ntap; ++t) {
        float* drow = p.dw + ((size_t)(tap0 + t) * p.Cout + co) * p.Cin + ci0;
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(trow + (uint32_t)(t * p.BN + c0), r);
          tmem_ld_wait();
          if (co < p.Cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (c0 + j < p.BN)
                red_add_v4(drow + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                           __uint_as_float(r[j + 3]));
          }
        }
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
static int g_last_wgrad_mode = 0;   // 0 staged-tile kernel, 1 tcgen05 (one x box per tap), 2 tcgen05 halo

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
    if (rc >= 0) { g_last_wgrad_mode = 0; return rc; }
  }
  for (int t = 0; t < d->ntaps; ++t)
    HWG_REQUIRE(d->tap_gy_h[t] == 0 && d->tap_gy_w[t] == 0,
                "hwg_conv_wgrad: per-tap gy phases are only supported for Cout, Cin in {16, 32}");
  PFN_encodeTiled encode = wg_get_encode();
  if (!encode) { set_error("hwg_conv_wgrad: cuTensorMapEncodeTiled unavailable"); return HWG_ERR_CUDA; }

  WgradKParams p;
  memset(&p, 0, sizeof(p));
  p.Cout = d->Cout; p.Cin = d->Cin; p.ntaps = d->ntaps;
  const int Hi = d->Hi > 0 ? d->Hi : d->Ho, Wi = d->Wi > 0 ? d->Wi : d->Wo;
  p.gsh = d->gy_stride_h > 1 ? d->gy_stride_h : 1;
  p.gsw = d->gy_stride_w > 1 ? d->gy_stride_w : 1;
  p.goh = d->gy_off_h; p.gow = d->gy_off_w;
  HWG_REQUIRE(p.gsh <= 4 && p.gsw <= 4, "hwg_conv_wgrad: gy strides up to 4");
  int dh_min = 1 << 30, dh_max = -(1 << 30), dw_min = 1 << 30, dw_max = -(1 << 30);
  for (int t = 0; t < d->ntaps; ++t) {
    p.tap_dh[t] = d->tap_dh[t]; p.tap_dw[t] = d->tap_dw[t];
    dh_min = d->tap_dh[t] < dh_min ? d->tap_dh[t] : dh_min; dh_max = d->tap_dh[t] > dh_max ? d->tap_dh[t] : dh_max;
    dw_min = d->tap_dw[t] < dw_min ? d->tap_dw[t] : dw_min; dw_max = d->tap_dw[t] > dw_max ? d->tap_dw[t] : dw_max;
  }
  p.dh_min = dh_min; p.dw_min = dw_min;
  p.co_tiles = (d->Cout + 127) / 128;
  p.CBa = d->Cout >= 64 ? 64 : d->Cout;
  p.CBb = d->Cin >= 64 ? 64 : d->Cin;
  p.nblk_a = ((d->Cout < 128 ? d->Cout : 128) + p.CBa - 1) / p.CBa;
  // Halo mode (HWG_WGRAD_HALO=1, off by default): ONE x box of 64 channels per chunk that covers every tap's shifted
  // window, the taps read it through MN-major descriptors that start at any 128-byte pixel row.  Numerically verified
  // on B200 (tests/test_conv_bwd_gpu.py under the switch) — the swizzle follows the absolute shared-memory address for
  // MN-major operands too — but SLOWER than one x box per tap: 310 vs 176 us on the 64->64 3x3 layer at 128 lines,
  // 1545 vs 652 us on the recognizer's conv1 (tools/wgrad_bench.py, profiles/README.md): a K step whose 16 pixel rows
  // do not start on an 8-row swizzle atom costs the tensor pipe about twice the shared-memory reads.
  static const bool halo_off = !(getenv("HWG_WGRAD_HALO") != nullptr && atoi(getenv("HWG_WGRAD_HALO")) == 1);
  const int span_h = dh_max - dh_min, span_w = dw_max - dw_min;
  const size_t ring_budget = 190 * 1024;
  bool configured = false;
  if (!halo_off && p.CBb == 64 && d->ntaps >= 2 && span_h <= 8 && span_w <= 16) {
    p.BN = d->Cin % 128 == 0 ? 128 : 64;
    const int nblk_b = p.BN / 64;
    int ph_top = 1;
    while (ph_top < 8 && ph_top < Hi) ph_top <<= 1;
    for (int pix = 128; pix >= 64 && !configured; pix >>= 1) {
      int PH = ph_top;
      while (pix / PH < 16) PH >>= 1;
      const int PW = pix / PH;
      if (pix == 128 && PW > 16 && PW / 2 >= Wi) continue;   // half of every tile would lie outside the row
      const int hb_w = PW + span_w, hb_h = PH + span_h;
      if (hb_w > 256 || hb_h > 256) continue;
      // one box per tap would stage ntaps * pix pixel rows: the halo must beat that by a margin
      if (hb_w * hb_h * 4 > d->ntaps * pix * 3) continue;
      const int hb_bytes = (hb_w * hb_h * 128 + 1023) / 1024 * 1024;
      const int a_bytes = p.nblk_a * pix * p.CBa * 2, b_bytes = nblk_b * hb_bytes;
      const int stages = (int)(ring_budget / (size_t)(a_bytes + b_bytes));
      if (stages < 2) continue;
      p.halo = 1; p.PW = PW; p.PH = PH; p.pix = pix; p.hb_w = hb_w; p.hb_h = hb_h; p.hb_bytes = hb_bytes;
      p.a_bytes = a_bytes; p.b_bytes = b_bytes; p.stages = stages > 6 ? 6 : stages;
      const int tpc_max = 512 / p.BN;
      p.tap_groups = (d->ntaps + tpc_max - 1) / tpc_max;
      p.tpc = (d->ntaps + p.tap_groups - 1) / p.tap_groups;
      configured = true;
    }
  }
  if (!configured) {
    // 64-pixel K chunks, as wide as the output row allows; one x box per tap of the group
    p.halo = 0; p.pix = 64;
    int PW = 64;
    while (PW > 8 && PW / 2 >= Wi) PW >>= 1;
    p.PW = PW; p.PH = p.pix / PW;
    p.BN = d->Cin % 256 == 0 ? 256 : (d->Cin % 128 == 0 ? 128 : (d->Cin >= 64 ? 64 : d->Cin));
    p.a_bytes = p.nblk_a * p.pix * p.CBa * 2;
    int tpc_max = 512 / p.BN;
    if (tpc_max > d->ntaps) tpc_max = d->ntaps;
    // at least three stages of {gy tile, tpc x boxes}
    while (tpc_max > 1 && (size_t)3 * (p.a_bytes + tpc_max * p.BN * p.pix * 2) > ring_budget) --tpc_max;
    p.tap_groups = (d->ntaps + tpc_max - 1) / tpc_max;
    p.tpc = (d->ntaps + p.tap_groups - 1) / p.tap_groups;
    p.b_bytes = p.tpc * p.BN * p.pix * 2;
    p.stages = (int)(ring_budget / (size_t)(p.a_bytes + p.b_bytes));
    if (p.stages > 8) p.stages = 8;
    HWG_REQUIRE(p.stages >= 2, "hwg_conv_wgrad: stage of %d bytes does not fit", p.a_bytes + p.b_bytes);
  }
  p.tap_groups = (d->ntaps + p.tpc - 1) / p.tpc;
  p.ci_tiles = d->Cin / p.BN;
  p.chunks_w = (Wi + p.PW - 1) / p.PW;
  p.chunks_h = (Hi + p.PH - 1) / p.PH;
  p.total_chunks = p.chunks_w * p.chunks_h * d->N;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  {
    const int cols = p.tpc * p.BN;
    p.tmem_cols = 32;
    while (p.tmem_cols < cols) p.tmem_cols <<= 1;
    HWG_REQUIRE(p.tmem_cols <= 512, "hwg_conv_wgrad: %d accumulator columns", cols);
  }
  p.dw = dw;
  const int units = p.tap_groups * p.co_tiles * p.ci_tiles;
  // split the pixel range so that the grid covers the SMs ~2x — or once when there is little work: every CTA ends with
  // 128 x BN x taps fp32 reductions into the arena, which at 16 lines per GPU cost as much as the operand stream
  // (B200, 16 lines: step 7.29 -> 7.09 ms with one wave)
  int cta_target = ((long long)p.total_chunks * (p.tap_groups * p.co_tiles * p.ci_tiles) >= 296LL * 16) ? 2 * 148 : 148;
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
  rc = p.halo ? encode_nhwc(encode, &tmx, x, d->Cin, d->W, d->H, d->N, d->x_pitch, p.CBb, p.hb_w, p.hb_h, 1, 1, "x halo")
              : encode_nhwc(encode, &tmx, x, d->Cin, d->W, d->H, d->N, d->x_pitch, p.CBb, p.PW, p.PH, 1, 1, "x");
  if (rc) return rc;
  // [barriers (1 KiB) | ring | one A tile of slack: a partially loaded gy tile is still addressed as 128 rows]
  const size_t smem = 1024 + 1024 + (size_t)p.stages * stage_bytes + (size_t)128 * p.pix * 2;
  HWG_REQUIRE(smem <= 227 * 1024, "hwg_conv_wgrad: %zu bytes of shared memory", smem);
  HWG_SMEM_OPTIN(conv_wgrad_kernel);
  p.unit_major = 1;
  if (const char* ov = getenv("HWG_WGRAD_ORDER")) p.unit_major = atoi(ov) != 0;   // development A/B switch
  if (splits > 65535) p.unit_major = 0;
  dim3 grid = p.unit_major ? dim3((unsigned)units, (unsigned)splits) : dim3((unsigned)splits, (unsigned)units);
  g_last_wgrad_mode = p.halo ? 2 : 1;
  conv_wgrad_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(tmg, tmx, p);
  return check_launch("conv_wgrad_kernel");
}

/* which kernel served the last hwg_conv_wgrad call of this process (tests / tools): 0 staged-tile, 1 tcgen05 with one
 * x box per tap, 2 tcgen05 with a halo box */
extern "C" int hwg_last_wgrad_kernel(void) { return g_last_wgrad_mode; }
