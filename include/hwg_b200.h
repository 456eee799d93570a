/*
 * hwg_b200.h — C-ABI of libhwg_b200.so, the sm_100a implementation of the
 * HWWithStyle training-step hot path of herobd/handwriting_line_generation.
 *
 * The reference has no FFI: its boundary is the PyTorch nn.Module surface
 * (SURVEY.md §8b).  Each entry point below names the reference call it stands
 * in for (file:line under /root/reference).  The Python package
 * handwriting_line_generation_b200 binds these with ctypes and wraps them in
 * torch.autograd.Functions behind modules whose constructor signatures and
 * state_dict keys are the reference's.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the library never allocates or frees device memory and keeps no pointer
 *    after a call returns: outputs, saved-for-backward tensors and workspaces
 *    are owned by the caller (PyTorch's caching allocator);
 *  - `stream` is a cudaStream_t (CUstream) passed as void*; all work is
 *    enqueued on it and nothing synchronises;
 *  - the return value is 0 on success, non-zero on error; hwg_last_error()
 *    returns the thread-local message.  No C++ exception crosses the ABI.
 */
#ifndef HWG_B200_H
#define HWG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HWG_OK 0
#define HWG_ERR_INVALID 1
#define HWG_ERR_CUDA 2
#define HWG_ERR_UNSUPPORTED 3

/* Library version (major*10000 + minor*100 + patch). */
int hwg_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* hwg_last_error(void);
/* Number of kernels this library has launched in this process (all threads).
 * bench.py reports the delta over its timed region as "gpu_launches". */
uint64_t hwg_launch_count(void);

/* ------------------------------------------------------------------------
 * CTC loss — replaces F.ctc_loss as called by CTCLoss, model/loss.py:28-30
 * (call sites trainer/hw_with_style_trainer.py:503,756,762).
 * blank = 0, log-space, fp32.  L = 2*S_max+1 is the row pitch of the
 * alpha/beta tables.
 *
 * log_probs   [T,B,C] fp32 contiguous log-softmax output
 * targets     int32, element (b,s) at targets[b*tgt_stride_b + s*tgt_stride_s]
 *             (the reference passes label.permute(1,0): a strided view)
 * input_lengths, target_lengths  int32 [B] on the device
 * ---------------------------------------------------------------------- */

/* Forward: alpha recursion (and, if log_beta != NULL, the beta recursion on a
 * second CTA per sequence, concurrently).  Writes nll[b] = -log p(target_b).
 * log_alpha/log_beta are [B,T,L] fp32 workspaces kept for the backward. */
int hwg_ctc_forward(const float* log_probs, int T, int B, int C,
                    const int32_t* targets, int64_t tgt_stride_b, int64_t tgt_stride_s,
                    int S_max, const int32_t* input_lengths,
                    const int32_t* target_lengths, int blank,
                    float* nll, float* log_alpha, float* log_beta, void* stream);

/* reduction='mean' of F.ctc_loss plus the reference wrapper's inf->0
 * (model/loss.py:30):  loss = mean_b(nll_b / max(S_b,1)), 0 if that is inf.
 * Also writes grad_nll_unit[b] = 1/(B*max(S_b,1)) (0 for every b if the loss
 * was inf) — the factor the backward multiplies by the incoming gradient. */
int hwg_ctc_reduce_mean(const float* nll, const int32_t* target_lengths, int B,
                        float* loss, float* grad_nll_unit, void* stream);

/* Backward: grad[t,b,c] = (exp(lp) - sum_{s:l'_s=c} exp(alpha+beta+nll-lp))
 *                         * grad_out[0] * grad_nll_unit[b]
 * and 0 for t >= input_lengths[b].  If beta_ready == 0 the beta recursion is
 * run here first (forward was called with log_beta == NULL). */
int hwg_ctc_backward(const float* grad_out, const float* grad_nll_unit,
                     const float* log_probs, int T, int B, int C,
                     const int32_t* targets, int64_t tgt_stride_b, int64_t tgt_stride_s,
                     int S_max, const int32_t* input_lengths,
                     const int32_t* target_lengths, int blank,
                     const float* nll, const float* log_alpha, float* log_beta,
                     int beta_ready, float* grad_log_probs, void* stream);

/* Best-path decode — replaces naive_decode, utils/string_utils.py:51-57
 * (caller getCER, trainer/hw_with_style_trainer.py:894-903): argmax over
 * classes (first maximum wins), drop repeats, drop blank.
 * raw      [T,B] int32 argmax per frame
 * decoded  [B,T] int32, first decoded_len[b] entries valid */
int hwg_ctc_greedy_decode(const float* log_probs, int T, int B, int C,
                          const int32_t* input_lengths, int blank,
                          int32_t* raw, int32_t* decoded, int32_t* decoded_len,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HWG_B200_H */
