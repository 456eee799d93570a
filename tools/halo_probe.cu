// One-tile experiment for the halo-tile convolution redesign — NOT part of the library.
//
// Question: can the A operand of tcgen05.mma read the taps of a 3x3 convolution as SHIFTED VIEWS of one halo tile in
// shared memory?  Output tile = 8 pixels wide x 16 high (M = 128 MMA rows, row m = ty*8 + tx); halo tile = 10 x 18 pixels
// of CK = 64 bf16 channels = 180 rows of 128 bytes, written the way TMA's SWIZZLE_128B writes a {64, 10, 18} box into a
// 1024-byte aligned buffer (16-byte chunk c of row r lands at chunk c ^ (r & 7), r = absolute row = address bits [7,10)).
// For tap (dy, dx) the K-major descriptor starts at base + (dy*10 + dx)*128 B with the 8-row-group stride SBO = 10*128 B,
// so that MMA row (ty, tx) reads halo pixel (ty + dy, tx + dx).  That start address is not 1024-byte aligned; whether the
// hardware swizzles by the ABSOLUTE shared-memory address (then the plain descriptor works) or relative to the start
// address corrected by the descriptor's `base_offset` field (bits [49,52) = (start >> 7) & 7) is what this measures.
//
// B = 64 x 64 identity (canonical K-major SWIZZLE_128B tile), so D[m][n] must equal halo[(ty+dy)*10 + tx+dx][n].
// Prints, for each descriptor variant and each of the nine taps, the number of mismatching outputs (0 = the view works).
//
// Build + run (on the B200 box):  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/halo_probe
//                                 tools/halo_probe.cu && gpurun_out/halo_probe      (tools/gpu_round2_first.sh does both)
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../handwriting_line_generation_b200/csrc/sm100.cuh"

using namespace hwg::sm100;

constexpr int HALO_W = 10, HALO_H = 18, HALO_PIX = HALO_W * HALO_H, CK = 64, NCH = 64;
constexpr int NVAR = 3;      // 0: plain descriptor, 1: base_offset = (start >> 7) & 7, 2: control (dy = dx = 0 only differs by start)

__host__ __device__ inline float halo_value(int pix, int ch) { return (float)((pix * 7 + ch * 3) % 251); }   // exact in bf16

__device__ __forceinline__ uint64_t desc_view(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                              // LBO: unused for swizzled K-major
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100)
  d |= (uint64_t)(base_offset & 7u) << 49;
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) halo_probe_kernel(float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                                  // 180 rows x 128 B (23040 B), padded to 24 KiB
  uint8_t* b_s = smem + 24 * 1024;                      // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // fill A and B with the 128-byte swizzle of an aligned TMA destination: chunk c of row r -> chunk c ^ (r & 7)
  for (int i = threadIdx.x; i < HALO_PIX * CK; i += 128) {
    const int r = i / CK, k = i % CK;
    const int chunk = (k >> 3) ^ (r & 7);
    reinterpret_cast<__nv_bfloat16*>(a_s + r * 128 + chunk * 16)[k & 7] = __float2bfloat16(halo_value(r, k));
  }
  for (int i = threadIdx.x; i < NCH * CK; i += 128) {
    const int n = i / CK, k = i % CK;
    const int chunk = (k >> 3) ^ (n & 7);
    reinterpret_cast<__nv_bfloat16*>(b_s + n * 128 + chunk * 16)[k & 7] = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
  const uint32_t idesc = hwg::sm100::umma_idesc_bf16(128, NCH);
  uint32_t phase = 0;

  for (int var = 0; var < NVAR; ++var) {
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
      if (threadIdx.x == 0) {
        uint32_t start = a_addr + (uint32_t)((dy * HALO_W + dx) * 128);
        uint32_t sbo = HALO_W * 128, bo = 0;
        if (var == 1) bo = (start >> 7) & 7u;
        if (var == 2) { start = a_addr + (uint32_t)(tap * 1024); sbo = 1024; }      // control: aligned canonical tiles
        const uint64_t da = desc_view(start, sbo, bo), db = desc_view(b_addr, 1024, 0);
#pragma unroll
        for (int kk = 0; kk < CK / 16; ++kk)            // K = 16 per MMA: +32 bytes inside the swizzle span
          umma_bf16(tmem_base, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, kk != 0 ? 1u : 0u);
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      float* o = out + (((size_t)var * 9 + tap) * 128 + warp * 32 + lane) * NCH;
      for (int c0 = 0; c0 < NCH; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o[c0 + j] = __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

int main() {
  const size_t n = (size_t)NVAR * 9 * 128 * NCH;
  float* d_out = nullptr;
  if (cudaMalloc(&d_out, n * sizeof(float)) != cudaSuccess) { printf("cudaMalloc failed\n"); return 2; }
  cudaMemset(d_out, 0xff, n * sizeof(float));
  const int smem = 24 * 1024 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  halo_probe_kernel<<<1, 128, smem>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
  float* h = (float*)malloc(n * sizeof(float));
  cudaMemcpy(h, d_out, n * sizeof(float), cudaMemcpyDeviceToHost);
  const char* names[NVAR] = {"shifted view, base_offset = 0            ", "shifted view, base_offset = (start>>7)&7 ",
                             "control: aligned tile at row 8*tap        "};
  int ok_var[NVAR] = {1, 1, 1};
  for (int var = 0; var < NVAR; ++var) {
    printf("%s:", names[var]);
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
      int bad = 0;
      for (int m = 0; m < 128; ++m) {
        const int ty = m / 8, tx = m % 8;
        const int pix = var == 2 ? tap * 8 + m : (ty + dy) * HALO_W + tx + dx;
        for (int c = 0; c < NCH; ++c) {
          const float want = pix < HALO_PIX ? halo_value(pix, c) : h[(((size_t)var * 9 + tap) * 128 + m) * NCH + c];
          bad += h[(((size_t)var * 9 + tap) * 128 + m) * NCH + c] != want;
        }
      }
      printf(" (%d,%d)=%d", dy, dx, bad);
      if (bad) ok_var[var] = 0;
    }
    printf("  -> %s\n", ok_var[var] ? "EXACT" : "mismatch");
  }
  printf("halo view usable: %s\n", ok_var[0] ? "yes, plain descriptor (swizzle follows the absolute address)"
                                   : (ok_var[1] ? "yes, with the base_offset field" : "NO"));
  return 0;
}
