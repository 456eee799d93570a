// Text spacing on the device (reference HWWithStyle.insert_spaces, model/hw_with_style.py:302-328).
//
// The reference walks batch x characters in Python: per character two `round(np.random.normal(counts[i,b,k].item(), std))`
// calls (2*L*B device->host synchronisations on a GPU), list concatenation, then a Python loop that sets the one-hot entries.
// Here the host draws the SAME standard normals from the same numpy stream in one vectorised call and hands them over;
//   spacing_plan_kernel  one block per line: count_i = max(0, rint(counts[i,b,0] + count_std * z)), dup_i likewise (double
//                        arithmetic without FMA contraction and round-half-to-even = what numpy + Python's round() do),
//                        exclusive prefix sums of count_i + dup_i -> offsets, the line length, ceil(max counts)
//   spacing_fill_kernel  spaced[t,b,:] = onehot(class of position t): binary search of t in the line's offsets
// The only host round trip left is ONE read of B+1 integers (the line lengths: the output's first dimension is data
// dependent and the generator's launch geometry needs it).  Integer work: bit-exact against the reference's goldens.
#include "common.cuh"

namespace hwg {
namespace {

constexpr int SP_THREADS = 256;

__device__ __forceinline__ long long load_label(const void* label, int is_i64, long long idx) {
  return is_i64 ? reinterpret_cast<const long long*>(label)[idx] : (long long)reinterpret_cast<const int*>(label)[idx];
}

__global__ void __launch_bounds__(SP_THREADS)
spacing_plan_kernel(const int32_t* __restrict__ lengths, const float* __restrict__ counts, int n_out,
                    const double* __restrict__ z, const long long* __restrict__ z_off, int L, int B, double count_std,
                    double dup_std, int32_t* __restrict__ reps, int32_t* __restrict__ offsets, int32_t* __restrict__ info) {
  __shared__ int scan[SP_THREADS];
  __shared__ int carry_s;
  __shared__ float wmax[SP_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  int len = lengths[b];
  len = len < 0 ? 0 : (len > L ? L : len);
  const double* zb = z + z_off[b];
  if (tid == 0) carry_s = 0;
  __syncthreads();
  // the reference's max_count looks at EVERY entry of counts (all positions, both channels): hw_with_style.py:303
  float mx = -3.4e38f;
  for (int i = tid; i < L * n_out; i += SP_THREADS) {
    const int ii = i / n_out, k = i - ii * n_out;
    mx = fmaxf(mx, counts[((size_t)ii * B + b) * n_out + k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) wmax[tid >> 5] = mx;
  for (int base = 0; base < len; base += SP_THREADS) {
    const int i = base + tid;
    int cnt = 0, dup = 0;
    if (i < len) {
      const float* c = counts + ((size_t)i * B + b) * n_out;
      // loc + scale * gauss in C doubles, two roundings (numpy legacy normal), then Python round() = half to even
      const double zc = zb[(size_t)i * n_out];
      long long r = (long long)rint(__dadd_rn((double)c[0], __dmul_rn(count_std, zc)));
      cnt = r < 0 ? 0 : (int)r;                           // [0] * negative == []
      if (n_out > 1) {
        const double zd = zb[(size_t)i * n_out + 1];
        r = (long long)rint(__dadd_rn((double)c[1], __dmul_rn(dup_std, zd)));
        dup = r < 0 ? 0 : (int)r;
      } else {
        dup = 1;
      }
      reps[((size_t)b * L + i) * 2] = cnt;
      reps[((size_t)b * L + i) * 2 + 1] = dup;
    }
    // inclusive scan of cnt + dup over the block
    scan[tid] = cnt + dup;
    __syncthreads();
    for (int o = 1; o < SP_THREADS; o <<= 1) {
      const int v = tid >= o ? scan[tid - o] : 0;
      __syncthreads();
      scan[tid] += v;
      __syncthreads();
    }
    const int carry = carry_s;
    if (i < len) offsets[(size_t)b * (L + 1) + i] = carry + scan[tid] - (cnt + dup);
    __syncthreads();
    if (tid == SP_THREADS - 1) carry_s = carry + scan[tid];
    __syncthreads();
  }
  if (tid == 0) {
    const int total = carry_s;
    for (int i = len; i <= L; ++i) offsets[(size_t)b * (L + 1) + i] = total;
    info[b] = total;
    float m = wmax[0];
    for (int w = 1; w < SP_THREADS / 32; ++w) m = fmaxf(m, wmax[w]);
    atomicMax(info + B, (int)ceilf(m));
  }
}

constexpr int FILL_ROWS = 8;

__global__ void __launch_bounds__(SP_THREADS)
spacing_fill_kernel(const void* __restrict__ label, int label_is_i64, long long ls_l, long long ls_b,
                    const int32_t* __restrict__ lengths, const int32_t* __restrict__ reps,
                    const int32_t* __restrict__ offsets, int L, int B, int T, int C, float* __restrict__ spaced) {
  __shared__ int cls_s[FILL_ROWS];
  const int b = blockIdx.y, t0 = blockIdx.x * FILL_ROWS, tid = threadIdx.x;
  int len = lengths[b];
  len = len < 0 ? 0 : (len > L ? L : len);
  const int32_t* off = offsets + (size_t)b * (L + 1);
  if (tid < FILL_ROWS) {
    const int t = t0 + tid;
    int cls = 0;                                          // behind the line: blank (hw_with_style.py:324-325)
    if (t < T && t < off[len]) {
      int lo = 0, hi = len - 1;                           // last character i with off[i] <= t
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= t) lo = mid; else hi = mid - 1;
      }
      const int cnt = reps[((size_t)b * L + lo) * 2];
      if (t - off[lo] >= cnt) {
        long long c = load_label(label, label_is_i64, (long long)lo * ls_l + (long long)b * ls_b);
        cls = c < 0 ? 0 : (c >= C ? C - 1 : (int)c);
      }
    }
    cls_s[tid] = cls;
  }
  __syncthreads();
  for (int e = tid; e < FILL_ROWS * C; e += SP_THREADS) {
    const int r = e / C, c = e - r * C, t = t0 + r;
    if (t < T) spaced[((size_t)t * B + b) * C + c] = (c == cls_s[r]) ? 1.f : 0.f;
  }
}

}  // namespace
}  // namespace hwg

using namespace hwg;

extern "C" int hwg_insert_spaces_plan(const int32_t* lengths, const float* counts, int n_out, const double* z,
                                      const int64_t* z_off, int L, int B, double count_std, double dup_std,
                                      int32_t* reps, int32_t* offsets, int32_t* info, void* stream) {
  HWG_REQUIRE(lengths && counts && z && z_off && reps && offsets && info, "hwg_insert_spaces_plan: null pointer");
  HWG_REQUIRE(L > 0 && B > 0 && (n_out == 1 || n_out == 2), "hwg_insert_spaces_plan: L=%d B=%d n_out=%d", L, B, n_out);
  cudaStream_t s = (cudaStream_t)stream;
  HWG_CUDA(cudaMemsetAsync(info + B, 0, sizeof(int32_t), s));
  spacing_plan_kernel<<<B, SP_THREADS, 0, s>>>(lengths, counts, n_out, z, reinterpret_cast<const long long*>(z_off), L, B,
                                               count_std, dup_std, reps, offsets, info);
  return check_launch("spacing_plan_kernel");
}

extern "C" int hwg_insert_spaces_fill(const void* label, int label_is_i64, int64_t label_stride_l, int64_t label_stride_b,
                                      const int32_t* lengths, const int32_t* reps, const int32_t* offsets, int L, int B,
                                      int T, int C, float* spaced, void* stream) {
  HWG_REQUIRE(label && lengths && reps && offsets && spaced, "hwg_insert_spaces_fill: null pointer");
  HWG_REQUIRE(L > 0 && B > 0 && T > 0 && C > 0 && B <= 65535, "hwg_insert_spaces_fill: L=%d B=%d T=%d C=%d", L, B, T, C);
  dim3 grid((unsigned)((T + FILL_ROWS - 1) / FILL_ROWS), (unsigned)B);
  spacing_fill_kernel<<<grid, SP_THREADS, 0, (cudaStream_t)stream>>>(label, label_is_i64, label_stride_l, label_stride_b,
                                                                     lengths, reps, offsets, L, B, T, C, spaced);
  return check_launch("spacing_fill_kernel");
}
