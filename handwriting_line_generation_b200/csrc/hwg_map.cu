// Batched small linear maps between parameter layouts and kernel layouts (one launch for a whole module).
//
// Every weight re-parameterisation of the generator is linear in the fp32 parameter and only mixes the kernel
// positions of one (row, column) channel pair: EqualLR scales (pure_gen.py:222-226), FusedUpsample's averaged 4x4
// kernel (pure_gen.py:259-271), the row sums of the nearest-upsample convolutions (pure_gen.py:176-186), the
// tap-major bf16 packing hwg_conv_fprop wants and its transpose for dgrad.  The adjoint of the same maps turns the
// tap-major fp32 output of hwg_conv_wgrad (and the per-channel sums of the fused backward kernels) into gradients in
// the parameters' own layouts.  A job is
//     dst[out_off[o] + r*d_r + c*d_c] (+)= sum_i M[i][o] * src[in_off(i) + r*s_r + c*s_c]      r < R, c < C
// with zeros written for R <= r < Rp and C <= c < Cp (operand padding).  This replaces the ~250 tiny
// permute/cat/pad/cast launches a training step otherwise spends on re-packing weights after every optimizer step.
#include "common.cuh"

namespace hwg {

constexpr int MAP_THREADS = 128, MAP_ITEMS = 2;   // (row, column) pairs per thread

__global__ void __launch_bounds__(MAP_THREADS) linear_map_kernel(const hwgMapJob* __restrict__ jobs,
                                                                 const int2* __restrict__ block_tab,
                                                                 const char* src_base, char* dst_base) {
  __shared__ hwgMapJob j;
  __shared__ float Ms[HWG_MAP_MAX * HWG_MAP_MAX];
  const int2 bt = block_tab[blockIdx.x];          // (job, first item / (MAP_THREADS*MAP_ITEMS))
  {
    const int* s = reinterpret_cast<const int*>(jobs + bt.x);
    int* d = reinterpret_cast<int*>(&j);
    for (int i = threadIdx.x; i < (int)(sizeof(hwgMapJob) / 4); i += blockDim.x) d[i] = s[i];
  }
  __syncthreads();
  const int nin = j.nin, nout = j.nout;
  const float jscale = j.scale * (j.scale_dev ? *j.scale_dev : 1.f);
  if (j.M != nullptr)
    for (int i = threadIdx.x; i < nin * nout; i += blockDim.x) Ms[i] = j.M[i];
  __syncthreads();
  const long long total = (long long)j.Rp * j.Cp;
  const char* sb = ((j.flags & 1) ? (const char*)nullptr : src_base) + j.src_off;
  char* dstb = ((j.flags & 2) ? (char*)nullptr : dst_base) + j.dst_off;
  for (int it = 0; it < MAP_ITEMS; ++it) {
    const long long idx = ((long long)bt.y * MAP_ITEMS + it) * MAP_THREADS + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx / j.Cp), c = (int)(idx % j.Cp);
    const bool pad = (r >= j.R) || (c >= j.C);
    const float* src = reinterpret_cast<const float*>(sb) + (long long)r * j.s_r + (long long)c * j.s_c;
    const long long drc = (long long)r * j.d_r + (long long)c * j.d_c;
    if (j.M == nullptr) {
      // all-ones column: dst = scale * sum_i src[i*in_stride]   (nin unbounded, nout == 1)
      float acc = 0.f;
      if (!pad)
        for (int i = 0; i < nin; ++i) acc += src[(long long)i * j.in_stride];
      acc *= jscale;
      const long long o = j.out_off[0] + drc;
      if (j.dst_bf16) reinterpret_cast<__nv_bfloat16*>(dstb)[o] = __float2bfloat16(acc);
      else if (j.accumulate) { if (!pad) reinterpret_cast<float*>(dstb)[o] += acc; }
      else reinterpret_cast<float*>(dstb)[o] = acc;
      continue;
    }
    float v[HWG_MAP_MAX];
#pragma unroll
    for (int i = 0; i < HWG_MAP_MAX; ++i) v[i] = (!pad && i < nin) ? src[j.in_off[i]] : 0.f;
    for (int o = 0; o < nout; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < HWG_MAP_MAX; ++i)
        if (i < nin) acc = fmaf(Ms[i * nout + o], v[i], acc);
      acc *= jscale;
      const long long oi = j.out_off[o] + drc;
      if (j.dst_bf16) reinterpret_cast<__nv_bfloat16*>(dstb)[oi] = __float2bfloat16(acc);
      else if (j.accumulate) { if (!pad) reinterpret_cast<float*>(dstb)[oi] += acc; }
      else reinterpret_cast<float*>(dstb)[oi] = acc;
    }
  }
}

}  // namespace hwg

extern "C" int hwg_map_items_per_block(void) { return hwg::MAP_THREADS * hwg::MAP_ITEMS; }

extern "C" int hwg_linear_map(const hwgMapJob* jobs_dev, const int32_t* block_tab_dev, int nblocks,
                              const void* src_base, void* dst_base, void* stream) {
  HWG_REQUIRE(jobs_dev && block_tab_dev && nblocks > 0, "hwg_linear_map: bad argument");
  hwg::linear_map_kernel<<<nblocks, hwg::MAP_THREADS, 0, (cudaStream_t)stream>>>(
      jobs_dev, reinterpret_cast<const int2*>(block_tab_dev), reinterpret_cast<const char*>(src_base),
      reinterpret_cast<char*>(dst_base));
  return hwg::check_launch("linear_map_kernel");
}
