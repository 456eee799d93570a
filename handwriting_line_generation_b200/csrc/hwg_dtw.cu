// DTW alignment of a label to the recognizer's output (reference model/hw_with_style.py:18-74 `correct_pred`; SURVEY.md §8
// row f3): the reference runs an O(T * (2S+1)) Python double loop of torch ops on the CPU per 'auto' lesson.  Here: one CTA
// per sequence, one thread per column j of the blank-interleaved label, the (i, j) recurrence as an anti-diagonal wavefront
// (cell (i, j) is computed at step i + j from the two previous diagonals kept in shared memory), the three-way argmin
// history as bytes in global memory, and the backtrack by one thread.  fp32, the same single subtraction and addition per
// cell as the reference, ties to the first of (up, diagonal, left) as torch.min does: the alignment is bit-exact.
// Bit-exact on the B200 against the reference's goldens and the CPU restatement (tests/test_dtw_gpu.py).
#include "common.cuh"
#include <math_constants.h>

namespace hwg {
namespace {

constexpr int DTW_MAX_L = 1024;        // columns (2 * label length + 1) one CTA can hold

// pred [T,B,C] fp32; label [S,B] int32 with element (s,b) at label[s*ls_s + b*ls_b]; hist [B,T,L] bytes;
// out [T+L, B] int32 (first out_len[b] rows valid for column b, the rest untouched), out_len [B].
__global__ void __launch_bounds__(DTW_MAX_L)
dtw_align_kernel(const float* __restrict__ pred, int T, int B, int C, const int* __restrict__ label, long long ls_s,
                 long long ls_b, int S, unsigned char* __restrict__ hist, int* __restrict__ out, int* __restrict__ out_len,
                 int* __restrict__ scratch /* [B, T+L] */) {
  extern __shared__ float sh[];
  const int L = 2 * S + 1;
  float* d1 = sh;                         // diagonal s-1, indexed by column 0..L
  float* d2 = sh + (L + 1);               // diagonal s-2
  int* lab = reinterpret_cast<int*>(sh + 2 * (L + 1));   // blank-interleaved label, columns 1..L -> lab[0..L-1]
  const int b = blockIdx.x, j = threadIdx.x;             // thread j owns column j (0..L)
  if (j < L) lab[j] = (j & 1) ? label[(long long)(j >> 1) * ls_s + (long long)b * ls_b] : 0;
  if (j <= L) { d1[j] = CUDART_INF_F; d2[j] = CUDART_INF_F; }
  __syncthreads();
  const int w = max(T / 2, abs(T - L));
  unsigned char* hb = hist + (size_t)b * T * L;
  const int cls = (j >= 1 && j <= L) ? lab[j - 1] : 0;
  for (int s = 0; s <= T + L; ++s) {
    const int i = s - j;
    float v = CUDART_INF_F;
    if (j <= L && i >= 0 && i <= T) {
      if (i == 0 && j == 0) {
        v = 0.f;
      } else if (i >= 1 && j >= 1 && j >= i - w && j <= i + w) {
        const float up = d1[j], left = d1[j - 1], diag = d2[j - 1];
        const float cost = 1.f - pred[((size_t)(i - 1) * B + b) * C + cls];
        int k = 0;
        float m = up;
        if (diag < m) { m = diag; k = 1; }
        if (left < m) { m = left; k = 2; }
        v = cost + m;
        hb[(size_t)(i - 1) * L + (j - 1)] = (unsigned char)k;
      }
    }
    __syncthreads();                       // everyone has read diagonals s-1 / s-2
    if (j <= L) { d2[j] = d1[j]; d1[j] = v; }
    __syncthreads();
  }
  if (j == 0) {
    // backtrack (hw_with_style.py:47-62): from (T-1, L-1) to (0, 0), collecting the label under the path
    int* rev = scratch + (size_t)b * (T + L);
    int i = T - 1, jj = L - 1, n = 0;
    rev[n++] = lab[jj];
    while (i > 0 || jj > 0) {
      const int h = hb[(size_t)i * L + jj];
      if (h == 0) { i -= 1; }
      else if (h == 1) { i -= 1; jj -= 1; }
      else { jj -= 1; }
      if (i < 0 || jj < 0) break;          // cannot happen for a history this kernel wrote; never read out of bounds
      rev[n++] = lab[jj];
    }
    for (int k = 0; k < n; ++k) out[(size_t)k * B + b] = rev[n - 1 - k];
    out_len[b] = n;
  }
}

}  // namespace
}  // namespace hwg

using namespace hwg;

extern "C" int hwg_dtw_align(const float* pred, int T, int B, int C, const int32_t* label, int64_t label_stride_s,
                             int64_t label_stride_b, int S, uint8_t* hist, int32_t* out, int32_t* out_len,
                             int32_t* scratch, void* stream) {
  HWG_REQUIRE(pred && label && hist && out && out_len && scratch, "hwg_dtw_align: null pointer");
  HWG_REQUIRE(T > 0 && B > 0 && C > 0 && S > 0, "hwg_dtw_align: empty extent (T=%d B=%d C=%d S=%d)", T, B, C, S);
  const int L = 2 * S + 1;
  HWG_REQUIRE(L + 1 <= DTW_MAX_L, "hwg_dtw_align: label length %d too long (2S+2 <= %d)", S, DTW_MAX_L);
  const int threads = ((L + 1 + 31) / 32) * 32;
  const size_t smem = (size_t)(2 * (L + 1)) * sizeof(float) + (size_t)L * sizeof(int);
  dtw_align_kernel<<<B, threads, smem, (cudaStream_t)stream>>>(pred, T, B, C, label, label_stride_s, label_stride_b, S, hist,
                                                            out, out_len, scratch);
  return check_launch("dtw_align_kernel");
}
