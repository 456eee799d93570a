// Gradient balancing over flat gradient buffers (reference trainer/hw_with_style_trainer.py:340-377,
// restated for the tests on the CPU).  The reference walks ~1300 parameter tensors per stashed gradient set with clone /
// abs / mean / `!= 0` host synchronisations; here the main gradient D and the K stashed sets R_k are flat fp32 buffers
// with the same parameter slots (optim.FlatAdam), and the step is three launches with no host round trip:
//   balance_stats : part[block][0] = sum|D|, part[block][1+k] = sum|R_k| over the block's chunk  (block table: (segment, chunk))
//   balance_coeffs: per segment the partial sums of its blocks added IN BLOCK ORDER (no atomics: every rank of a
//                   data-parallel job, given the same all-reduced gradients, computes the same bits), mean|.|, the zero-mean
//                   replacement (:354-359), mult[k][seg] = x_k * mean|D| / mean|R_k|
//   balance_apply : D[i] += sum_k mult[k][seg(i)] * R_k[i]
// Parity: tests/test_balance_gpu.py (against the CPU restatement pinned to the unmodified trainer).
#include "common.cuh"

namespace hwg {
namespace {

constexpr int BAL_THREADS = 256;
constexpr int BAL_CHUNK = 4096;          // elements of one segment handled by one block
constexpr int BAL_MAX_SETS = 8;

struct BalSets { const float* r[BAL_MAX_SETS]; };

// block_tab[b] = (segment, chunk index within the segment)
__global__ void __launch_bounds__(BAL_THREADS)
balance_stats_kernel(const float* __restrict__ d, BalSets sets, int K, const long long* __restrict__ seg_off,
                     const long long* __restrict__ seg_len, const int2* __restrict__ block_tab, float* __restrict__ part,
                     int* __restrict__ seg_first) {
  __shared__ float red[BAL_THREADS / 32][BAL_MAX_SETS + 1];
  const int2 bt = block_tab[blockIdx.x];
  const long long base = seg_off[bt.x], len = seg_len[bt.x];
  const long long lo = (long long)bt.y * BAL_CHUNK, hi = lo + BAL_CHUNK < len ? lo + BAL_CHUNK : len;
  float acc[BAL_MAX_SETS + 1];
#pragma unroll
  for (int k = 0; k <= BAL_MAX_SETS; ++k) acc[k] = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += BAL_THREADS) {
    acc[0] += fabsf(d[base + i]);
#pragma unroll
    for (int k = 0; k < BAL_MAX_SETS; ++k)
      if (k < K) acc[1 + k] += fabsf(sets.r[k][base + i]);
  }
#pragma unroll
  for (int k = 0; k <= BAL_MAX_SETS; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x <= K) {
    float t = 0.f;
    for (int w = 0; w < BAL_THREADS / 32; ++w) t += red[w][threadIdx.x];
    part[(size_t)blockIdx.x * (K + 1) + threadIdx.x] = t;
  }
  if (threadIdx.x == 0 && bt.y == 0) seg_first[bt.x] = (int)blockIdx.x;   // the blocks of a segment are consecutive rows
}

// one block: means, the fill value (average of the non-zero mean|D|), multipliers
__global__ void __launch_bounds__(BAL_THREADS)
balance_coeffs_kernel(const float* __restrict__ part, const int* __restrict__ seg_first,
                      const long long* __restrict__ seg_len, int nseg, int K,
                      BalSets xs /* r[0] = device pointer to x[K] */, float* __restrict__ mult) {
  __shared__ float red_s[BAL_THREADS / 32], red_c[BAL_THREADS / 32];
  __shared__ float fill_s;
  // ordered sum of the partials of segment s, column k
  auto seg_sum = [&](int s, int k) {
    const int nb = (int)((seg_len[s] + BAL_CHUNK - 1) / BAL_CHUNK), f = seg_first[s];
    float t = 0.f;
    for (int c = 0; c < nb; ++c) t += part[(size_t)(f + c) * (K + 1) + k];
    return t;
  };
  float nz_sum = 0.f, nz_cnt = 0.f;
  for (int s = threadIdx.x; s < nseg; s += BAL_THREADS) {
    const float m = seg_sum(s, 0) / (float)seg_len[s];
    if (m != 0.f) { nz_sum += m; nz_cnt += 1.f; }
  }
  nz_sum = warp_sum(nz_sum);
  nz_cnt = warp_sum(nz_cnt);
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = nz_sum; red_c[threadIdx.x >> 5] = nz_cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < BAL_THREADS / 32; ++w) { a += red_s[w]; c += red_c[w]; }
    fill_s = c > 0.f ? a / c : 0.f;
  }
  __syncthreads();
  const float fill = fill_s;
  const float* x = xs.r[0];
  for (int s = threadIdx.x; s < nseg; s += BAL_THREADS) {
    const float n = (float)seg_len[s];
    float md = seg_sum(s, 0) / n;
    if (md == 0.f) md = fill;                                  // :354-359
    for (int k = 0; k < K; ++k) {
      const float mr = seg_sum(s, 1 + k) / n;
      mult[(size_t)k * nseg + s] = mr != 0.f ? x[k] * (md / mr) : 0.f;     // :373-376 (`if abmean_R != 0`)
    }
  }
}

__global__ void __launch_bounds__(BAL_THREADS)
balance_apply_kernel(float* __restrict__ d, BalSets sets, int K, const long long* __restrict__ seg_off,
                     const long long* __restrict__ seg_len, const int2* __restrict__ block_tab, int nseg,
                     const float* __restrict__ mult) {
  const int2 bt = block_tab[blockIdx.x];
  const long long base = seg_off[bt.x], len = seg_len[bt.x];
  const long long lo = (long long)bt.y * BAL_CHUNK, hi = lo + BAL_CHUNK < len ? lo + BAL_CHUNK : len;
  float m[BAL_MAX_SETS];
#pragma unroll
  for (int k = 0; k < BAL_MAX_SETS; ++k) m[k] = k < K ? mult[(size_t)k * nseg + bt.x] : 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += BAL_THREADS) {
    float v = d[base + i];
    // the reference adds the sets one after the other (:369-376): same order, same roundings
#pragma unroll
    for (int k = 0; k < BAL_MAX_SETS; ++k)
      if (k < K && m[k] != 0.f) v += m[k] * sets.r[k][base + i];
    d[base + i] = v;
  }
}

}  // namespace
}  // namespace hwg

using namespace hwg;

extern "C" int hwg_balance_chunk(void) { return BAL_CHUNK; }

extern "C" int hwg_balance(float* g_main, const float* const* sets_host, int K, const float* x_dev,
                           const int64_t* seg_off_dev, const int64_t* seg_len_dev, int nseg,
                           const int32_t* block_tab_dev, int nblocks, float* sums_dev, float* mult_dev, void* stream) {
  HWG_REQUIRE(g_main && sets_host && x_dev && seg_off_dev && seg_len_dev && block_tab_dev && sums_dev && mult_dev,
              "hwg_balance: null pointer");
  HWG_REQUIRE(K >= 1 && K <= BAL_MAX_SETS && nseg > 0 && nblocks > 0, "hwg_balance: K=%d (1..%d), nseg=%d, nblocks=%d", K,
              BAL_MAX_SETS, nseg, nblocks);
  BalSets sets{}, xs{};
  for (int k = 0; k < K; ++k) {
    HWG_REQUIRE(sets_host[k] != nullptr, "hwg_balance: set %d is null", k);
    sets.r[k] = sets_host[k];
  }
  xs.r[0] = x_dev;
  cudaStream_t s = (cudaStream_t)stream;
  const long long* off = reinterpret_cast<const long long*>(seg_off_dev);
  const long long* len = reinterpret_cast<const long long*>(seg_len_dev);
  const int2* tab = reinterpret_cast<const int2*>(block_tab_dev);
  int* seg_first = reinterpret_cast<int*>(sums_dev + (size_t)nblocks * (K + 1));
  balance_stats_kernel<<<nblocks, BAL_THREADS, 0, s>>>(g_main, sets, K, off, len, tab, sums_dev, seg_first);
  if (int rc = check_launch("balance_stats_kernel")) return rc;
  balance_coeffs_kernel<<<1, BAL_THREADS, 0, s>>>(sums_dev, seg_first, len, nseg, K, xs, mult_dev);
  if (int rc = check_launch("balance_coeffs_kernel")) return rc;
  balance_apply_kernel<<<nblocks, BAL_THREADS, 0, s>>>(g_main, sets, K, off, len, tab, nseg, mult_dev);
  return check_launch("balance_apply_kernel");
}
