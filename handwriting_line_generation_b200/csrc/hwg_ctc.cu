// CTC loss forward/backward and best-path decode for sm_100a.
//
// Stands in for F.ctc_loss as called from CTCLoss (reference model/loss.py:28-30)
// and for naive_decode (reference utils/string_utils.py:51-57).
//
// Layout in HBM
//   log_probs  [T,B,C] fp32, C contiguous (what the drop-in recognizer head stores)
//   log_alpha  [B,T,L] fp32, L = 2*S_max+1 (only s < 2*S_b+1, t < T_b are touched)
//   log_beta   [B,T,L] fp32
//   grad       [T,B,C] fp32
//
// Kernels
//   ctc_chain_kernel   one CTA per (sequence, direction).  The alpha and the beta
//                      recursions are independent dependency chains of T_b steps,
//                      so they run on two CTAs at the same time.  One thread per
//                      augmented-label state; the previous column lives in a
//                      double-buffered shared array padded with -inf so that the
//                      three-way log-sum-exp has no boundary branches.  The
//                      log-prob rows of the next TC frames are staged into shared
//                      memory with cp.async (16/8/4-byte, chosen from C's
//                      alignment) one chunk ahead of the recursion, so the chain
//                      never waits on HBM.
//   ctc_grad_kernel    one CTA per (sequence, block of frames): per-state
//                      occupancies exp(alpha+beta+nll-lp) are summed per class
//                      through a per-sequence CSR (class -> label positions) built
//                      once in shared memory: deterministic, no atomics.  Reads
//                      log_probs and writes grad fully coalesced along C.
//   ctc_reduce_mean_kernel, ctc_decode_kernel  (tiny)
#include "common.cuh"
#include <math_constants.h>

namespace hwg {

__device__ __forceinline__ void cp_async_f(float* smem_dst, const float* gsrc, int nfloat_vec) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (nfloat_vec == 4)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
  else if (nfloat_vec == 2)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// log(exp(a)+exp(b)+exp(c)) with the -inf convention of ATen's LossCTC.cpp
// (max == -inf -> treat max as 0, giving log(0) = -inf).
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return logf(expf(a - m) + expf(b - m) + expf(c - m)) + m;
}

struct ChainSmem {
  int* lab;      // [Lpad] augmented labels l'
  float* col;    // [2][Lpad+4] previous/current column, 2 pads of -inf each side
  float* lp;     // [2][TC][Cp] staged log-prob rows
};

// Labels index the staged log-prob row and the per-class tables: a value outside [0, C) (invalid input — ATen's CPU path
// raises, the host wrapper checks CPU targets) is clamped so that it can never address memory out of range.
__device__ __forceinline__ int clamp_label(int l, int C) { return l < 0 ? 0 : (l >= C ? C - 1 : l); }

template <int VEC>
__global__ void __launch_bounds__(1024)
ctc_chain_kernel(const float* __restrict__ lp, int T, int B, int C,
                 const int32_t* __restrict__ targets, int64_t ts_b, int64_t ts_s, int S_max,
                 const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len,
                 int blank, float* __restrict__ nll, float* __restrict__ log_alpha,
                 float* __restrict__ log_beta, int dir_base, int TC, int Cp) {
  const int b = blockIdx.x;
  const int dir = blockIdx.y + dir_base;  // 0: alpha (t ascending), 1: beta (t descending)
  const int tid = threadIdx.x, NT = blockDim.x;
  const int L = 2 * S_max + 1;
  int Tb = in_len[b]; Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int Sb = tg_len[b]; Sb = Sb < 0 ? 0 : (Sb > S_max ? S_max : Sb);
  const int Lb = 2 * Sb + 1;
  const int colpitch = L + 4;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* lps = reinterpret_cast<float*>(smem_raw);              // [2][TC][Cp]
  float* col = lps + 2 * TC * Cp;                               // [2][colpitch]
  int* lab = reinterpret_cast<int*>(col + 2 * colpitch);        // [L]

  for (int s = tid; s < Lb; s += NT)
    lab[s] = (s & 1) ? clamp_label(targets[b * ts_b + (int64_t)(s >> 1) * ts_s], C) : blank;
  for (int i = tid; i < 2 * colpitch; i += NT) col[i] = -CUDART_INF_F;

  float* table = (dir == 0 ? log_alpha : log_beta) + (size_t)b * T * L;
  const int nchunks = (Tb + TC - 1) / TC;
  const int vec_per_row = C / VEC;

  auto issue_chunk = [&](int k) {
    float* dst = lps + (k & 1) * TC * Cp;
    int j0 = k * TC;
    int rows = min(TC, Tb - j0);
    for (int i = tid; i < rows * vec_per_row; i += NT) {
      int r = i / vec_per_row, v = i - r * vec_per_row;
      int t = dir == 0 ? (j0 + r) : (Tb - 1 - (j0 + r));
      cp_async_f(dst + r * Cp + v * VEC, lp + ((size_t)t * B + b) * C + v * VEC, VEC);
    }
    cp_async_commit();
  };

  if (nchunks > 0) issue_chunk(0);
  __syncthreads();  // lab / col initialised

  for (int k = 0; k < nchunks; ++k) {
    if (k + 1 < nchunks) { issue_chunk(k + 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const float* rows = lps + (k & 1) * TC * Cp;
    const int j0 = k * TC;
    const int jn = min(TC, Tb - j0);
    for (int jj = 0; jj < jn; ++jj) {
      const int j = j0 + jj;
      const int t = dir == 0 ? j : (Tb - 1 - j);
      const float* row = rows + jj * Cp;
      const float* prev = col + ((j + 1) & 1) * colpitch + 2;
      float* cur = col + (j & 1) * colpitch + 2;
      for (int s = tid; s < Lb; s += NT) {
        const int l = lab[s];
        float v;
        if (j == 0) {
          if (dir == 0) v = (s <= 1) ? row[l] : -CUDART_INF_F;
          else v = (s >= Lb - 2) ? row[l] : -CUDART_INF_F;
        } else if (dir == 0) {
          float a1 = prev[s], a2 = prev[s - 1];
          float a3 = (s >= 2 && lab[s - 2] != l) ? prev[s - 2] : -CUDART_INF_F;
          v = lse3(a1, a2, a3) + row[l];
        } else {
          float b1 = prev[s], b2 = prev[s + 1];
          float b3 = (s + 2 < Lb && lab[s + 2] != l) ? prev[s + 2] : -CUDART_INF_F;
          // prev[Lb], prev[Lb+1] are the -inf pads (never written: s < Lb)
          v = lse3(b1, b2, b3) + row[l];
        }
        cur[s] = v;
        table[(size_t)t * L + s] = v;
      }
      __syncthreads();
    }
  }

  if (dir == 0 && tid == 0) {
    float r = CUDART_INF_F;
    if (Tb > 0) {
      const float* last = col + ((Tb - 1) & 1) * colpitch + 2;
      float l1 = last[Lb - 1];
      float l2 = Lb > 1 ? last[Lb - 2] : -CUDART_INF_F;
      float m = fmaxf(l1, l2);
      r = (m == -CUDART_INF_F) ? CUDART_INF_F : -(logf(expf(l1 - m) + expf(l2 - m)) + m);
    }
    nll[b] = r;
  }
}

// loss = mean_b(nll_b / max(S_b,1)); inf -> 0 (reference model/loss.py:30).
__global__ void ctc_reduce_mean_kernel(const float* __restrict__ nll,
                                       const int32_t* __restrict__ tg_len, int B,
                                       float* __restrict__ loss, float* __restrict__ unit) {
  // single warp, fixed order: deterministic
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) {
    int s = tg_len[b]; s = s < 1 ? 1 : s;
    acc += nll[b] / (float)s;
  }
  acc = warp_sum(acc);
  float mean = acc / (float)B;
  bool bad = isinf(mean);
  if (threadIdx.x == 0) loss[0] = bad ? 0.f : mean;
  for (int b = threadIdx.x; b < B; b += 32) {
    int s = tg_len[b]; s = s < 1 ? 1 : s;
    unit[b] = bad ? 0.f : 1.f / ((float)B * (float)s);
  }
}

constexpr int GRAD_TT = 8;  // frames per barrier pair: one warp per frame in the class-sum phase

__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ grad_out, const float* __restrict__ unit,
                const float* __restrict__ lp, int T, int B, int C,
                const int32_t* __restrict__ targets, int64_t ts_b, int64_t ts_s, int S_max,
                const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len,
                int blank, const float* __restrict__ nll, const float* __restrict__ log_alpha,
                const float* __restrict__ log_beta, float* __restrict__ grad, int frames_per_cta) {
  const int b = blockIdx.y;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int L = 2 * S_max + 1;
  int Tb = in_len[b]; Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int Sb = tg_len[b]; Sb = Sb < 0 ? 0 : (Sb > S_max ? S_max : Sb);
  const int Lb = 2 * Sb + 1;
  const int t_begin = blockIdx.x * frames_per_cta;
  const int t_end = min(T, t_begin + frames_per_cta);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* lab = reinterpret_cast<int*>(smem_raw);       // [S_max] labels
  int* cls_off = lab + S_max;                        // [C+1]
  int* cls_pos = cls_off + (C + 1);                  // [S_max] label positions grouped by class
  float* occ = reinterpret_cast<float*>(cls_pos + S_max);  // [GRAD_TT][L]

  for (int i = tid; i < Sb; i += NT) lab[i] = clamp_label(targets[b * ts_b + (int64_t)i * ts_s], C);
  __syncthreads();
  // counting sort of label positions by class (stable -> fixed summation order)
  for (int c = tid; c < C; c += NT) {
    int n = 0;
    for (int i = 0; i < Sb; ++i) n += (lab[i] == c);
    cls_off[c + 1] = n;
  }
  if (tid == 0) cls_off[0] = 0;
  __syncthreads();
  if (tid == 0) for (int c = 0; c < C; ++c) cls_off[c + 1] += cls_off[c];
  __syncthreads();
  for (int c = tid; c < C; c += NT) {
    int o = cls_off[c];
    for (int i = 0; i < Sb; ++i) if (lab[i] == c) cls_pos[o++] = i;
  }
  __syncthreads();

  const float nll_b = nll[b];
  const float scale = grad_out[0] * unit[b];
  const float* A = log_alpha + (size_t)b * T * L;
  const float* Bt = log_beta + (size_t)b * T * L;
  const int warp = tid >> 5, lane = tid & 31, nwarp = NT >> 5;

  for (int t0 = t_begin; t0 < t_end; t0 += GRAD_TT) {
    const int nt = min(GRAD_TT, t_end - t0);
    // phase 1: per-state occupancy gamma_t(s) = exp(alpha+beta+nll-lp)
    for (int i = tid; i < nt * Lb; i += NT) {
      int tt = i / Lb, s = i - tt * Lb, t = t0 + tt;
      float g = 0.f;
      if (t < Tb) {
        int l = (s & 1) ? lab[s >> 1] : blank;
        float ab = A[(size_t)t * L + s] + Bt[(size_t)t * L + s];
        g = expf(ab + nll_b - lp[((size_t)t * B + b) * C + l]);
      }
      occ[tt * L + s] = g;
    }
    __syncthreads();
    // phase 2: per-class sums, one warp per frame, lanes strided over classes
    for (int tt = warp; tt < nt; tt += nwarp) {
      const int t = t0 + tt;
      const float* o = occ + tt * L;
      float* grow = grad + ((size_t)t * B + b) * C;
      const float* lrow = lp + ((size_t)t * B + b) * C;
      if (t >= Tb) {
        for (int c = lane; c < C; c += 32) grow[c] = 0.f;
        continue;
      }
      // blank: all even states, reduced in a fixed order
      float bsum = 0.f;
      for (int i = lane; i <= Sb; i += 32) bsum += o[2 * i];
      bsum = warp_sum(bsum);
      for (int c = lane; c < C; c += 32) {
        float acc = (c == blank) ? bsum : 0.f;
        for (int q = cls_off[c]; q < cls_off[c + 1]; ++q) acc += o[2 * cls_pos[q] + 1];
        grow[c] = (expf(lrow[c]) - acc) * scale;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
ctc_decode_kernel(const float* __restrict__ lp, int T, int B, int C,
                  const int32_t* __restrict__ in_len, int blank, int32_t* __restrict__ raw,
                  int32_t* __restrict__ decoded, int32_t* __restrict__ decoded_len) {
  extern __shared__ int raws[];  // [T]
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  int Tb = in_len ? in_len[b] : T; Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  for (int t = warp; t < T; t += nwarp) {
    const float* row = lp + ((size_t)t * B + b) * C;
    float best = -CUDART_INF_F; int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      float v = row[c];
      if (v != v) v = CUDART_INF_F;  // NaN sorts as the maximum (numpy argmax)
      if (bi == 0x7fffffff || v > best) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) { raws[t] = bi; raw[(size_t)t * B + b] = bi; }
  }
  __syncthreads();
  if (warp == 0) {
    int count = 0;
    for (int base = 0; base < Tb; base += 32) {
      int t = base + lane;
      int r = t < Tb ? raws[t] : blank;
      bool keep = t < Tb && r != blank && (t == 0 || r != raws[t - 1]);
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) decoded[(size_t)b * T + count + __popc(m & ((1u << lane) - 1u))] = r;
      count += __popc(m);
    }
    if (lane == 0) decoded_len[b] = count;
  }
}

static int chain_launch(const float* lp, int T, int B, int C, const int32_t* targets,
                        int64_t ts_b, int64_t ts_s, int S_max, const int32_t* in_len,
                        const int32_t* tg_len, int blank, float* nll, float* la, float* lb,
                        int dir_base, int ndir, cudaStream_t st) {
  const int L = 2 * S_max + 1;
  int NT = ((L + 31) / 32) * 32;
  if (NT > 1024) NT = 1024;
  int vec = (C % 4 == 0) ? 4 : (C % 2 == 0 ? 2 : 1);
  if ((reinterpret_cast<uintptr_t>(lp) & 15) != 0) vec = 1;
  int Cp = ((C + 3) / 4) * 4;
  int TC = 6144 / Cp; TC = TC < 1 ? 1 : (TC > 32 ? 32 : TC);
  size_t smem = (size_t)2 * TC * Cp * 4 + (size_t)2 * (L + 4) * 4 + (size_t)L * 4;
  dim3 grid(B, ndir);
#define LAUNCH_CHAIN(V)                                                                      \
  do {                                                                                       \
    HWG_SMEM_OPTIN(ctc_chain_kernel<V>);                                                     \
    ctc_chain_kernel<V><<<grid, NT, smem, st>>>(lp, T, B, C, targets, ts_b, ts_s, S_max,     \
                                                in_len, tg_len, blank, nll, la, lb,          \
                                                dir_base, TC, Cp);                           \
  } while (0)
  HWG_REQUIRE(smem <= 200 * 1024, "ctc: C=%d, S_max=%d need %zu B of shared memory", C, S_max, smem);
  if (vec == 4) LAUNCH_CHAIN(4);
  else if (vec == 2) LAUNCH_CHAIN(2);
  else LAUNCH_CHAIN(1);
#undef LAUNCH_CHAIN
  return check_launch("ctc_chain_kernel");
}

}  // namespace hwg

using namespace hwg;

extern "C" int hwg_ctc_forward(const float* log_probs, int T, int B, int C,
                               const int32_t* targets, int64_t tgt_stride_b,
                               int64_t tgt_stride_s, int S_max, const int32_t* input_lengths,
                               const int32_t* target_lengths, int blank, float* nll,
                               float* log_alpha, float* log_beta, void* stream) {
  HWG_REQUIRE(log_probs && nll && log_alpha && input_lengths && target_lengths,
              "hwg_ctc_forward: null pointer");
  HWG_REQUIRE(T > 0 && B > 0 && C > 0 && S_max >= 0, "hwg_ctc_forward: bad shape T=%d B=%d C=%d S=%d", T, B, C, S_max);
  HWG_REQUIRE(S_max == 0 || targets, "hwg_ctc_forward: null targets");
  HWG_REQUIRE(blank >= 0 && blank < C, "hwg_ctc_forward: blank %d out of range", blank);
  return chain_launch(log_probs, T, B, C, targets, tgt_stride_b, tgt_stride_s, S_max,
                      input_lengths, target_lengths, blank, nll, log_alpha, log_beta, 0,
                      log_beta ? 2 : 1, (cudaStream_t)stream);
}

extern "C" int hwg_ctc_reduce_mean(const float* nll, const int32_t* target_lengths, int B,
                                   float* loss, float* grad_nll_unit, void* stream) {
  HWG_REQUIRE(nll && target_lengths && loss && grad_nll_unit && B > 0, "hwg_ctc_reduce_mean: bad argument");
  ctc_reduce_mean_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(nll, target_lengths, B, loss, grad_nll_unit);
  return check_launch("ctc_reduce_mean_kernel");
}

extern "C" int hwg_ctc_backward(const float* grad_out, const float* grad_nll_unit,
                                const float* log_probs, int T, int B, int C,
                                const int32_t* targets, int64_t tgt_stride_b,
                                int64_t tgt_stride_s, int S_max, const int32_t* input_lengths,
                                const int32_t* target_lengths, int blank, const float* nll,
                                const float* log_alpha, float* log_beta, int beta_ready,
                                float* grad_log_probs, void* stream) {
  HWG_REQUIRE(grad_out && grad_nll_unit && log_probs && nll && log_alpha && log_beta && grad_log_probs,
              "hwg_ctc_backward: null pointer");
  HWG_REQUIRE(T > 0 && B > 0 && C > 0 && S_max >= 0, "hwg_ctc_backward: bad shape");
  HWG_REQUIRE(blank >= 0 && blank < C, "hwg_ctc_backward: blank %d out of range", blank);
  cudaStream_t st = (cudaStream_t)stream;
  if (!beta_ready) {
    int rc = chain_launch(log_probs, T, B, C, targets, tgt_stride_b, tgt_stride_s, S_max,
                          input_lengths, target_lengths, blank, nullptr, nullptr, log_beta, 1, 1, st);
    if (rc) return rc;
  }
  const int L = 2 * S_max + 1;
  int frames = 16;
  // enough CTAs to cover the SMs a few times, but at least GRAD_TT*2 frames each
  while (frames > 8 && (long)((T + frames - 1) / frames) * B < 2 * 148) frames >>= 1;
  dim3 grid((T + frames - 1) / frames, B);
  size_t smem = (size_t)(2 * S_max + C + 1) * 4 + (size_t)GRAD_TT * L * 4;
  HWG_REQUIRE(smem <= 200 * 1024, "ctc backward: needs %zu B of shared memory", smem);
  HWG_SMEM_OPTIN(ctc_grad_kernel);
  ctc_grad_kernel<<<grid, 256, smem, st>>>(grad_out, grad_nll_unit, log_probs, T, B, C, targets,
                                           tgt_stride_b, tgt_stride_s, S_max, input_lengths,
                                           target_lengths, blank, nll, log_alpha, log_beta,
                                           grad_log_probs, frames);
  return check_launch("ctc_grad_kernel");
}

extern "C" int hwg_ctc_greedy_decode(const float* log_probs, int T, int B, int C,
                                     const int32_t* input_lengths, int blank, int32_t* raw,
                                     int32_t* decoded, int32_t* decoded_len, void* stream) {
  HWG_REQUIRE(log_probs && raw && decoded && decoded_len, "hwg_ctc_greedy_decode: null pointer");
  HWG_REQUIRE(T > 0 && B > 0 && C > 0, "hwg_ctc_greedy_decode: bad shape");
  size_t smem = (size_t)T * 4;
  HWG_REQUIRE(smem <= 200 * 1024, "hwg_ctc_greedy_decode: T=%d too long", T);
  HWG_SMEM_OPTIN(ctc_decode_kernel);
  ctc_decode_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(log_probs, T, B, C, input_lengths, blank,
                                                            raw, decoded, decoded_len);
  return check_launch("ctc_decode_kernel");
}
