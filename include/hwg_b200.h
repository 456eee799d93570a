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


/* ------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (bf16 x bf16 -> fp32 in
 * TMEM), TMA-staged NHWC tiles.  One entry point covers every dense
 * convolution of the hot path, because a launch is described by a list of
 * taps (input-pixel offsets) rather than by kernel/stride/padding:
 *
 *   y[n, ho, wo, co] = epilogue( sum_t sum_ci x[n, ho+dh[t], wo+dw[t], ci]
 *                                             * w[t][co][ci] )
 *
 * with x read as zero outside [0,H)x[0,W) (TMA out-of-bounds fill = the
 * reference's zero padding).  Replaces, in the reference:
 *   nn.Conv2d 3x3 (pure_gen.py:181-183,197; cnn_only_hwr.py:31) : 9 taps
 *   nn.Conv1d k3 dilated (cnn_only_hwr.py:78-90)                : 3 taps, H=1
 *   nn.Upsample(2,1)+Conv2d (pure_gen.py:176-186)   : 2 launches (row parity),
 *       6 taps each, weights pre-summed over the duplicated rows
 *   FusedUpsample conv_transpose2d 4x4 s2 (pure_gen.py:259-279) : 4 launches
 *       (output parity), 4 taps each
 *   nn.ConvTranspose2d (4,3) on H=1 (pure_gen.py:161-163)        : 4 launches
 *       (output row), 3 taps each
 * The y strides/pointer select the output rows/columns a launch writes.
 *
 * x        bf16 NHWC, channel pitch x_pitch (elements, multiple of 8)
 * w        bf16 [ntaps][Cout][Cin] (Cin contiguous), Cin multiple of 16
 * y        bf16 or fp32; element (n,ho,wo,c) at y[n*y_stride_n + ho*y_stride_h
 *          + wo*y_stride_w + c]
 * Epilogue, in this order: + bias[co]; + noise_w[co]*noise[n,ho,wo,co];
 * activation; accumulate per-(n,co) sum / sum of squares into stats; store.
 * HWG_EPI_LOGSOFTMAX replaces the activation by log-softmax over co
 * (cnn_only_hwr.py:91 + the permute at :105 via the y strides); needs
 * Cout <= 256.
 * ---------------------------------------------------------------------- */
#define HWG_MAX_TAPS 16
#define HWG_ACT_NONE 0
#define HWG_ACT_RELU 1
#define HWG_ACT_LRELU 2
#define HWG_ACT_LOGSOFTMAX 3
#define HWG_DT_BF16 0
#define HWG_DT_F32 1

typedef struct hwgConvDesc {
  int32_t N, H, W;        /* input extent */
  int32_t Cin;            /* channels contracted (multiple of 16) */
  int32_t x_pitch;        /* channel pitch of x, elements */
  int32_t Cout;           /* output channels */
  int32_t Ho, Wo;         /* output grid computed by this launch */
  int32_t ntaps;
  int32_t tap_dh[HWG_MAX_TAPS];
  int32_t tap_dw[HWG_MAX_TAPS];
  int64_t y_stride_n, y_stride_h, y_stride_w; /* elements */
  int32_t y_dtype;        /* HWG_DT_* */
  int32_t act;            /* HWG_ACT_* */
  float slope;            /* LeakyReLU negative slope */
  int32_t tile_w;         /* 0 = auto; else output-tile width (8..128, power of 2) */
  int64_t nz_stride_n, nz_stride_h, nz_stride_w; /* noise tensor strides, elements */
  int32_t in_stride_h, in_stride_w; /* 0/1 = dense; s>1: tap t reads x[n, ho*s_h+dh[t], wo*s_w+dw[t], :]
                                       (strided convolution = the input gradient of a stride-s transposed conv) */
  uint64_t noise_seed;    /* in-kernel N(0,1) noise (counter-based hash + Box-Muller) when noise == NULL and noise_w != NULL */
  uint64_t noise_subseq;  /* distinguishes launches that share a seed */
  uint64_t noise_seed_dev; /* 0, or the DEVICE address of a uint64 that is added to noise_seed when the kernel
                              runs: lets a captured CUDA graph draw fresh noise on every replay */
  /* Channel folding (0 = off): Cout = F * fold_c; output channel f*fold_c + ch is channel ch of the output pixel
   * displaced by (f / fold_w) * fold_stride_h + (f % fold_w) * fold_stride_w elements.  Runs launches that share
   * input and taps as one: the 4 output rows of ConvTranspose2d (4,3) (pure_gen.py:161-163) and the 4 output
   * parities of FusedUpsample (pure_gen.py:259-279; taps = the 3x3 union, unused (tap, parity) weights zero).
   * bias / noise_w are [Cout] (replicated per fold by the caller); stats is [N][fold_c][2]; fold f draws its
   * in-kernel noise from subsequence noise_subseq + f. */
  int32_t fold_c, fold_w;
  int64_t fold_stride_h, fold_stride_w;
  /* 0: every tap feeds every fold (w is [ntaps][Cout][Cin]).  k > 0: taps [f*k, (f+1)*k) belong to fold f only and
   * w is [ntaps][fold_c][Cin] — a stride-2 transposed convolution as one launch whose folds are the output
   * parities (4 taps each).  k > 0 is served by the staged-tile kernel (hwg_conv_small.cu; Cin in {16,32,64}). */
  int32_t fold_taps;
  int32_t force_tcgen05;  /* 1: never route to the staged-tile kernel (benchmark / test switch) */
} hwgConvDesc;

/* bias [Cout] fp32 or NULL; noise_w [Cout] fp32 or NULL (no noise); noise fp32 tensor or NULL
 * (NULL with noise_w set: draw N(0,1) in the kernel, keyed by noise_seed and the element index);
 * stats [N][Cout][2] fp32 (sum, sum of squares; accumulated with atomics, the
 * caller zeroes it) or NULL. */
int hwg_conv_fprop(const hwgConvDesc* desc, const void* x, const void* w, const float* bias,
                   const float* noise, const float* noise_w, float* stats, void* y,
                   void* stream);
/* Which kernel served the most recent hwg_conv_fprop call: 1 = conv_fprop_kernel (tcgen05 implicit GEMM, the
 * tensor-bound layers), 2 = conv_small_kernel (TMA-staged tiles + mma.sync, the HBM-bound 16-64 channel layers),
 * 3 / 4 = conv_fprop_kernel with the halo-tile main loop (development switch HWG_CONV_HALO; 3 = one 2-D halo tile per
 * K chunk, 4 = one halo tile per kernel row).  For benchmarks / profiles. */
int hwg_last_conv_kernel(void);

/* ------------------------------------------------------------------------
 * Fused memory-bound passes around the convolutions.  Activations are NHWC
 * bf16 with C a multiple of 8; `HW` below is pixels per image.
 * ---------------------------------------------------------------------- */

/* out[b,o] = act(sum_k in[b,k]*W[o,k] + bias[o]) in fp32 — the style MLP
 * Linear+LeakyReLU(0.2) layers (pure_gen.py:31-39) and the AdaIN style
 * projections (pure_gen.py:57,63), all ten concatenated into one call. */
int hwg_linear_f32(const float* in, const float* W, const float* bias, float* out,
                   int B, int K, int O, int act, float slope, void* stream);

/* PixelNorm (pure_gen.py:306-311): out = in / sqrt(mean_k(in^2) + 1e-8), [B,K] fp32. */
int hwg_pixelnorm_f32(const float* in, float* out, int B, int K, void* stream);

/* Generator input (pure_gen.py:43-48): content [T,B,C] fp32 (element (t,b,c) at
 * content[t*cs_t + b*cs_b + c*cs_c]) and style [B,S] fp32 -> x [B,1,T,Cp] bf16 with
 * channels [0,C) = content, [C,C+S) = style broadcast over t, rest 0. */
int hwg_gen_pack_input(const float* content, int64_t cs_t, int64_t cs_b, int64_t cs_c,
                       const float* style, int T, int B, int C, int S, int Cp, void* x,
                       void* stream);

/* AdaIN coefficients (pure_gen.py:62-69): from per-(n,c) sum/sumsq over HW pixels
 * (biased variance, eps) and gamma/beta rows: coef[n,c] = (a, b) with
 * a = gamma*rstd, b = beta - mean*a, so that AdaIN(x) = a*x + b. */
int hwg_adain_coeffs(const float* stats, const float* gamma, const float* beta,
                     int64_t gb_stride_n, int N, int C, int HW, float eps, float* coef,
                     float* save_mean_rstd /* [N,C,2] or NULL: kept for the backward */,
                     void* stream);

/* BatchNorm coefficients (nn.BatchNorm2d/1d in cnn_only_hwr.py:36,79): training
 * (use_batch_stats=1): reduce stats over n, biased variance to normalise,
 * running_mean/var updated with `momentum` and the unbiased variance;
 * eval: running stats.  coef[c] = (a, b). */
int hwg_bn_coeffs(const float* stats, int N, int C, int64_t count_per_n,
                  const float* weight, const float* bias, float* running_mean,
                  float* running_var, float momentum, float eps, int use_batch_stats,
                  float* coef, float* save_mean_rstd, void* stream);

/* y = act(a*x + b) on NHWC bf16; coef is [N,C,2] (per_sample=1) or [C,2]. In place allowed. */
int hwg_scale_shift_act(const void* x, void* y, const float* coef, int per_sample, int N,
                        int64_t HW, int C, int act, float slope, void* stream);

/* Blur 3x3 [1,2,1]x[1,2,1]/16 with zero padding (pure_gen.py:80-137) fused with
 * NoiseInjection (:72-79), LeakyReLU and the InstanceNorm statistics of the
 * result.  x, y [N,H,W,C] bf16; noise fp32 NHWC or NULL (NULL + noise_w: in-kernel RNG). */
int hwg_blur_noise_act_stats(const void* x, void* y, int N, int H, int W, int C,
                             const float* noise, const float* noise_w, uint64_t noise_seed,
                             uint64_t noise_subseq, const uint64_t* noise_seed_dev /* or NULL */,
                             int act, float slope, float* stats, void* stream);

/* Generator output (pure_gen.py:29,50): AdaIN apply (coef [N,C,2]) + 1x1 conv C->1
 * (weight w[C], bias *b0) + tanh; x [N,H,W,C] bf16 -> out [N,1,H,W] fp32. */
int hwg_gen_output(const void* x, const float* coef, const float* w, const float* b0 /* device, 1 float */,
                   int N, int64_t HW, int C, float* out, void* stream);

/* Recognizer stem (cnn_only_hwr.py:44-46): Conv2d(1,64,3,pad 1)+ReLU+MaxPool2d(2,2)
 * in one pass.  img [N,1,H,W] fp32 -> y [N,H/2,W/2,Cout] bf16; w [Cout,9], b [Cout]. */
int hwg_hwr_stem(const float* img, const float* w, const float* b, int N, int H, int W,
                 int Cout, void* y, void* stream);

/* MaxPool2d on NHWC bf16 with -inf padding (cnn_only_hwr.py:48,51-52,55-56). */
int hwg_maxpool_nhwc(const void* x, void* y, int N, int H, int W, int C, int kh, int kw,
                     int sh, int sw, int ph, int pw, int Ho, int Wo, void* stream);

/* ------------------------------------------------------------------------
 * Backward of the convolutions.
 * dgrad needs no new entry point: it is hwg_conv_fprop on the output gradient with the
 * negated taps and transposed weight matrices (the host packs them).
 *
 * wgrad (replaces cudnnConvolutionBackwardFilter behind nn.Conv2d/Conv1d backward):
 *   dw[t][co][ci] += sum_{n,ho,wo} gy[n,ho,wo,co] * x[n, ho+dh[t], wo+dw[t], ci]
 * as a tcgen05 GEMM with M = Cout tile (128), N = Cin tile (<= 256), K = pixels, both
 * operands MN-major straight from the NHWC tensors via TMA (zero fill outside x = padding),
 * split over pixel ranges with fp32 vector reductions (red.global.add.v4.f32) into dw.
 * x  bf16 NHWC [N,H,W,x_pitch], Cin in {16, 32} or a multiple of 64;  gy bf16 NHWC [N,Ho,Wo,gy_pitch],
 * Cout in {16, 32} or a multiple of 8 >= 64
 * dw fp32 [ntaps][Cout][Cin], ACCUMULATED into (caller zeroes it).
 * ---------------------------------------------------------------------- */
typedef struct hwgWgradDesc {
  int32_t N, H, W, Cin, x_pitch;
  int32_t Ho, Wo, Cout, gy_pitch;
  int32_t ntaps;
  int32_t tap_dh[HWG_MAX_TAPS];
  int32_t tap_dw[HWG_MAX_TAPS];
  /* Iteration grid (0 = Ho x Wo) and where grid point (i,j) sits in gy: gy[n, i*gy_stride_h+gy_off_h,
   * j*gy_stride_w+gy_off_w, :] pairs with x[n, i+dh, j+dw, :].  Strides > 1 give the weight gradient of one
   * output phase of an up-sampling convolution (pure_gen.py:176-186, 259-279). */
  int32_t Hi, Wi, gy_stride_h, gy_stride_w, gy_off_h, gy_off_w;
  /* Per-tap gy phase (added to gy_off; 0 <= phase < stride): all output phases of an up-sampling convolution in
   * ONE launch, so that gy is read once.  Supported for Cout, Cin in {16, 32} (the staged-tile kernel
   * wgrad_small_kernel: one TMA box of gy and one halo box of x per spatial tile, ldmatrix + mma.sync over all
   * taps, accumulators resident in registers; HBM-bound layers). */
  int32_t tap_gy_h[HWG_MAX_TAPS];
  int32_t tap_gy_w[HWG_MAX_TAPS];
} hwgWgradDesc;

int hwg_conv_wgrad(const hwgWgradDesc* desc, const void* x, const void* gy, float* dw, void* stream);
/* Which kernel served the most recent hwg_conv_wgrad call: 0 = wgrad_small_kernel (staged tiles + mma.sync, Cout and
 * Cin in {16, 32}), 1 = conv_wgrad_kernel (tcgen05, split-K, the gy tile of a pixel chunk staged once for a group of
 * taps, one x box per tap), 2 = conv_wgrad_kernel in halo mode (one x box per chunk that covers every tap's shifted
 * window; Cin a multiple of 64; development switch HWG_WGRAD_HALO=1, off by default: verified but measured slower).  For benchmarks / profiles. */
int hwg_last_wgrad_kernel(void);

/* ------------------------------------------------------------------------
 * Memory-bound backward passes of the recognizer (NHWC bf16 gradients).
 * Each pass also emits the per-channel sum of the gradient it writes = the
 * bias gradient of the convolution that produced that activation.
 * ---------------------------------------------------------------------- */

/* LogSoftmax backward (cnn_only_hwr.py:91): gz = g - exp(lp) * sum_c g, written as the NHWC bf16
 * tensor [B,1,T,Cp] the head's dgrad/wgrad read (channels >= C zero).  g, lp are [T,B,C] fp32.
 * dbias [C] fp32 is accumulated (caller zeroes). */
int hwg_logsoftmax_bwd(const float* g, const float* lp, int T, int B, int C, int Cp, void* gz,
                       float* dbias, void* stream);

/* BatchNorm(+ReLU) backward, pass 1: sums[c] = (sum gy, sum gy*xhat), gy = g * (a*z+b > 0),
 * xhat = (z-mean)*rstd.  g, z [rows,C] bf16; coef [C,2] = forward (a,b); save [C,2] = (mean,rstd).
 * sums [C,2] fp32 accumulated (caller zeroes). */
int hwg_bn_bwd_reduce(const void* g, const void* z, const float* coef, const float* save, int64_t rows,
                      int C, int relu, float* sums, void* stream);
/* pass 2: gz = weight*rstd*(gy - sums0/M - xhat*sums1/M) -> bf16 [rows,C]; dweight = sums1, dbias = sums0
 * are read by the host from `sums`; dconv_bias [C] fp32 accumulates sum gz (caller zeroes).
 * M = norm_rows if > 0, else rows: with statistics synchronised over a process group (SyncBN) `sums` holds the
 * all-reduced pair and norm_rows the global row count, while `rows` stays the local extent of g / z / gz. */
int hwg_bn_bwd_apply(const void* g, const void* z, const float* coef, const float* save, const float* weight,
                     const float* sums, int64_t rows, int64_t norm_rows, int C, int relu, void* gz,
                     float* dconv_bias, void* stream);

/* ReLU + MaxPool2d backward in gather form (no atomics): gc[n,h,w,:] = (c>0) * sum over the pooling
 * windows that contain (h,w) and whose first maximum is (h,w) of ga[window].  c [N,H,W,C] is the
 * saved post-ReLU pre-pool activation, ga [N,Ho,Wo,C].  dbias [C] accumulates sum gc. */
int hwg_relu_maxpool_bwd(const void* ga, const void* c, int N, int H, int W, int C, int kh, int kw, int sh,
                         int sw, int ph, int pw, int Ho, int Wo, void* gc, float* dbias, void* stream);

/* Stem backward (conv0+ReLU+MaxPool, cnn_only_hwr.py:44-46): dw [Cout,9], db [Cout] fp32 accumulated
 * from ga [N,H/2,W/2,Cout] bf16 and the fp32 image (pre-activations are recomputed). */
int hwg_hwr_stem_bwd(const float* img, const float* w, const float* b, const void* ga, int N, int H, int W,
                     int Cout, float* dw, float* db, void* stream);
/* d loss / d image of the stem in ONE pass (the GAN lessons back-propagate through the recognizer's input,
 * trainer/hw_with_style_trainer.py:752-764): conv0 1->64 3x3 pad 1 + ReLU + MaxPool 2x2 (cnn_only_hwr.py:44-46).
 * img [N,1,H,W] fp32, w [64,9], b [64] fp32, ga [N,H/2,W/2,64] bf16 = gradient w.r.t. the pooled output.
 * gimg [N,1,H,W] fp32 is ADDED to (caller zeroes).  Equivalent to hwg_hwr_stem_bwd_expand followed by the
 * 9-tap transposed convolution with conv0's weights, without materialising the [N,H,W,64] gradient. */
int hwg_hwr_stem_bwd_image(const float* img, const float* w, const float* b, const void* ga, int N, int H,
                           int W, int Cout, float* gimg, void* stream);

/* Same recomputation, but writes the gradient w.r.t. the conv0 OUTPUT, gc0 [N,H,W,Cout] bf16 (ga routed to the
 * arg-max of each 2x2 window where it passed the ReLU, zero elsewhere).  The image gradient the GAN lessons
 * need (trainer/hw_with_style_trainer.py:760-764: the generated line is recognised and the CTC loss flows
 * back into the generator) is then hwg_conv_fprop(gc0) with the 9 transposed taps of conv0. */
int hwg_hwr_stem_bwd_expand(const float* img, const float* w, const float* b, const void* ga, int N, int H,
                            int W, int Cout, void* gc0, void* stream);

/* ------------------------------------------------------------------------
 * Memory-bound backward passes of the generator.
 * Half-block forward (pure_gen.py:205-214):  y = conv(+blur) + nw*z ; a = LeakyReLU(y) ;
 * x_next = gamma*IN(a)+beta = A*a + B  with A = gamma*rstd.
 * ---------------------------------------------------------------------- */

/* pass 1: sums[n,c] = (sum_hw g, sum_hw g*ahat), ahat = (a-mean)*rstd.  g, a [N,H,W,C] bf16;
 * save [N,C,2] = (mean, rstd).  sums [N,C,2] accumulated (caller zeroes): sums[...,0] = dbeta,
 * sums[...,1] = dgamma. */
int hwg_adain_bwd_reduce(const void* g, const void* a, const float* save, int N, int64_t HW, int C,
                         float* sums, void* stream);
/* pass 2: ga = A*(g - s0/HW - ahat*s1/HW); gy = ga * (a > 0 ? 1 : slope) -> bf16 [N,H,W,C];
 * dch[c] = (sum gy  [= conv bias gradient when no blur follows],  sum gy*z [= noise weight gradient]),
 * accumulated (caller zeroes).  z is the forward's noise: `noise` tensor (fp32 NHWC) or regenerated from
 * (noise_seed, noise_subseq); row_subseq != 0 reproduces the per-output-row launches of the initial
 * transposed conv (subsequence = noise_subseq + h, element index without h). */
int hwg_adain_bwd_apply(const void* g, const void* a, const float* save, const float* coef,
                        const float* sums, int N, int H, int W, int C, float slope,
                        const float* noise, uint64_t noise_seed, uint64_t noise_subseq,
                        const uint64_t* noise_seed_dev /* or NULL */, int row_subseq,
                        void* gy, float* dch, void* stream);

/* Generator output backward (pure_gen.py:29,50): out = tanh(sum_c w[c]*(A*a+B)[c] + b0).
 * g_out, out [N,1,H,W] fp32; a [N,H,W,C] bf16; coef [N,C,2] = (A,B).  Writes gx [N,H,W,C] bf16 = gradient
 * w.r.t. the (virtual) AdaIN output; accumulates dw[c] and db0 (caller zeroes; dwb = [C+1] floats). */
int hwg_gen_output_bwd(const float* g_out, const float* out, const void* a, const float* coef,
                       const float* w, int N, int64_t HW, int C, void* gx, float* dwb, void* stream);

/* ------------------------------------------------------------------------
 * Batched linear maps between parameter layouts and kernel layouts — one launch
 * for every weight of a module.  Replaces, in the reference, the per-forward
 * weight re-parameterisations done with ATen ops: the EqualLR pre-hook
 * (pure_gen.py:222-226,239-241), FusedUpsample.forward's pad/average
 * (pure_gen.py:259-271), and — together with their adjoints — what autograd
 * records for them; it also produces the tap-major bf16 operands of
 * hwg_conv_fprop (forward and dgrad) and maps hwg_conv_wgrad's tap-major fp32
 * output back to the parameters' own layouts.
 *
 *   dst[out_off[o] + r*d_r + c*d_c] (+)= scale * sum_i M[i*nout + o] * src[in_off[i] + r*s_r + c*s_c]
 *
 * for r < R, c < C; zeros are written for R <= r < Rp or C <= c < Cp (operand
 * padding; skipped when accumulating).  With M == NULL the job is a plain sum
 * over nin strided inputs (in_off[i] = i*in_stride, nout = 1) — batch
 * reductions of per-sample sums.  src is fp32; dst fp32 or bf16.
 * Offsets/strides are in ELEMENTS of the respective dtype except src_off/dst_off
 * (BYTES, added to the src_base/dst_base launch arguments unless the job's flags
 * mark them as absolute device addresses).  The job table and the M tables live in device memory.
 * ---------------------------------------------------------------------- */
#define HWG_MAP_MAX 16
typedef struct hwgMapJob {
  int64_t src_off, dst_off;        /* bytes, relative to src_base / dst_base */
  const float* M;                  /* [nin][nout] device pointer, or NULL (sum of strided inputs) */
  int32_t nin, nout;               /* nin, nout <= HWG_MAP_MAX unless M == NULL */
  int32_t R, C, Rp, Cp;
  int32_t dst_bf16, accumulate;
  float scale; int32_t flags;      /* bit 0: src_off is an absolute address; bit 1: dst_off is */
  int64_t in_stride;               /* used when M == NULL */
  int64_t s_r, s_c, d_r, d_c;
  int64_t in_off[HWG_MAP_MAX], out_off[HWG_MAP_MAX];
  const float* scale_dev;          /* optional device scalar multiplied into `scale` at run time (spectral norm 1/sigma) */
} hwgMapJob;
/* block_tab_dev: nblocks pairs (job index, block index within the job) of int32 in device memory; one block covers
 * hwg_map_items_per_block() consecutive (r, c) items of its job, so a job needs ceil(Rp*Cp / that) blocks. */
int hwg_map_items_per_block(void);
int hwg_linear_map(const hwgMapJob* jobs_dev, const int32_t* block_tab_dev, int nblocks,
                   const void* src_base, void* dst_base, void* stream);

/* ------------------------------------------------------------------------
 * Flat fused optimizer step — replaces clip_grad_value_ + torch.optim.Adam.step
 * + zero_grad of the reference trainer (trainer/hw_with_style_trainer.py:381-391,
 * base/base_trainer.py:95-100; Adam lr 2e-4, betas (0.5, 0.999), weight_decay 0)
 * for an optimizer group stored as ONE flat fp32 buffer (n elements, multiple
 * of 4, 16-byte aligned): p, g, exp_avg m, exp_avg_sq v.
 *   g' = clamp(g*grad_scale, +-clip_value) (clip_value <= 0: no clipping)
 *   m = lerp(m, g', 1-beta1); v = beta2*v + (1-beta2)*g'^2
 *   p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps)
 * t = ++step_dev[0] (device-resident float counter, so a captured launch advances
 * on every CUDA-graph replay).  zero_grad != 0 clears g on the way out. */
int hwg_adam_flat(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float clip_value, float grad_scale, float* step_dev,
                  int zero_grad, void* stream);

/* Backward of y = act(x W^T + b), fp32, act in {none, LeakyReLU(slope)} — the style MLP (pure_gen.py:31-39) and the
 * concatenated AdaIN projections (pure_gen.py:57,63); replaces autograd's addmm/leaky_relu backward kernels.
 * x [B,K], y [B,O] (post-activation; only read for LeakyReLU), gy [B,O], W [O,K].
 * gx [B,K] or NULL (ADDED to: the caller zeroes it); gW [O,K] or NULL, gb [O] or NULL (written, or added to when accumulate != 0 — e.g.
 * straight into the flat gradient buffer of hwg_adam_flat). */
int hwg_linear_bwd_f32(const float* x, const float* y, const float* gy, const float* W, int B, int K,
                       int O, int act, float slope, float* gx, float* gW, float* gb, int accumulate,
                       void* stream);

/* ------------------------------------------------------------------------
 * Gradient balancing on flat gradient buffers — SURVEY.md 8 f2, reference
 * trainer/hw_with_style_trainer.py:340-377 (`balance_loss`), restated in oracle/balance.py:
 *   for every stashed set R_k and parameter segment s:  D_s += x_k * R_k,s * (mean|D_s| / mean|R_k,s|)
 * with mean|D_s| taken before any set is added, segments whose mean|D_s| is exactly 0 using the average of the
 * non-zero means instead (:354-359), and sets with mean|R_k,s| == 0 skipped (:373).
 * g_main and every set are flat fp32 buffers with the same parameter slots (optim.FlatAdam): seg_off_dev[s] /
 * seg_len_dev[s] = first element / element count of parameter s; block_tab_dev = nblocks (segment, chunk) int32
 * pairs, one per hwg_balance_chunk() elements of a segment, the rows of a segment consecutive and in chunk order;
 * sets_host = HOST array of K device pointers (K <= 8); x_dev = the K multipliers (`balance_var_x`) in device memory;
 * sums_dev (nblocks*(K+1) + nseg floats: per-block partial sums + the first block of every segment) and mult_dev
 * [K][nseg] are workspaces.  Three launches, no host synchronisation, no atomics: the result is a deterministic function
 * of the inputs, so data-parallel ranks holding the same all-reduced gradients stay bit-identical. */
int hwg_balance_chunk(void);
int hwg_balance(float* g_main, const float* const* sets_host, int K, const float* x_dev,
                const int64_t* seg_off_dev, const int64_t* seg_len_dev, int nseg, const int32_t* block_tab_dev,
                int nblocks, float* sums_dev, float* mult_dev, void* stream);

/* ------------------------------------------------------------------------
 * Discriminator (reference model/discriminator_ap.py:68-161):
 * memory-bound passes around the tensor-core convolutions.  NHWC bf16, C % 8 == 0.
 * ---------------------------------------------------------------------- */
/* in_conv (nn.Conv2d(1, dim, 7, padding=(0,3)), :75) as an implicit GEMM: out[n,h,w,j] = img[n,0,h,w+j-pad]
 * (zero outside the row) for j < kw, 0 for kw <= j < 16 -> bf16 [N,H,W,16]; the 7x7 convolution is then the 7-tap
 * (dy, 0) hwg_conv_fprop over these 16 channels.  hwg_shift_collapse is the adjoint (image gradient from the
 * dgrad of that convolution): dimg[n,0,h,x] (+)= sum_j g[n,h,x-j+pad,j]. */
int hwg_shift_expand(const float* img, void* out, int N, int H, int W, int kw, int pad, void* stream);

/* One-input-channel stem convolution straight from the fp32 image: the 7x7 in_conv of DiscriminatorAP
 * (model/discriminator_ap.py:75, padding (0,3)) and the 5x5 down_conv1[0] of Encoder2 (model/autoencoder.py:345, padding 2),
 * forward only.  img [N,1,H,W] fp32 (rounded to bf16 on the way in, like hwg_shift_expand does); w = the tap-major bf16
 * operand of the shift-expansion route, [kh][Cout][16] with column j = kernel column j (columns >= kw are ignored);
 * bias [Cout] fp32 or NULL; y [N,Ho,Wo,Cout] bf16 NHWC contiguous, Ho = H + 2*pad_h - kh + 1, Wo = W + 2*pad_w - kw + 1
 * (zero padding); stats [N][Cout][2] fp32 (sum, sum of squares of y before rounding; accumulated, caller zeroes) or
 * NULL.  kh, kw <= 8, Cout 32 or 64.  Write-bound (2*Cout bytes per pixel): staged image tile + mma.sync im2col fragments.
 * Same result as hwg_shift_expand + a kh-tap hwg_conv_fprop up to the fp32 summation order. */
int hwg_stem_conv(const float* img, const void* w, const float* bias, int N, int H, int W, int kh, int kw,
                  int pad_h, int pad_w, int Cout, void* y, float* stats, void* stream);
int hwg_shift_collapse(const void* g, float* dimg, int N, int H, int W, int kw, int pad, int accumulate,
                       void* stream);
/* nn.GroupNorm(groups, C) (:76,101) from the per-(n,c) sums of the conv epilogue: coef[n,c] = (a, b) with
 * GroupNorm(z) = a*z + b (apply with hwg_scale_shift_act, per_sample=1, LeakyReLU(0.1));
 * save[n,c] = (mean, rstd) of c's group (biased variance, eps). */
int hwg_gn_coeffs(const float* stats, const float* gamma, const float* beta, int N, int C, int groups,
                  int64_t HW, float eps, float* coef, float* save_mean_rstd, void* stream);
/* nn.AvgPool2d((kh,kw)) (stride = kernel, floor): [N,H,W,C] -> [N,H/kh,W/kw,C]. */
int hwg_avgpool_nhwc(const void* x, void* y, int N, int H, int W, int C, int kh, int kw, void* stream);
/* Backward of y = LeakyReLU(scale[n,c]*conv) [-> AvgPool2d(kh,kw)] (Dropout2d channel scale, :89-90 etc.):
 * gz[n,h,w,c] = scale[n,c] * (y > 0 ? 1 : slope) * g[n,h/kh,w/kw,c] / (kh*kw); y is the stored post-activation
 * tensor [N,H,W,C], g [N,H/kh,W/kw,C]; scale [N,C] fp32 or NULL (= 1). */
int hwg_act_bwd(const void* g, const void* y, const float* scale, float slope, int N, int H, int W, int C, int kh,
                int kw, void* gz, void* stream);
/* Backward of a = LeakyReLU(GroupNorm(z)) [-> AvgPool2d(kh,kw)], three launches:
 *   hwg_norm_bwd_reduce: sums[n,c] += (sum gy', sum gy'*z), gy' = g/(kh*kw) * (a*z+b > 0 ? 1 : slope)   (sums zeroed by the caller)
 *   hwg_gn_bwd_coeffs  : spq[n,c] = (sc, P, Q) from the group sums; dgamma[c] += sum gy'*xhat, dbeta[c] += sum gy' (or NULL)
 *   hwg_norm_bwd_apply : gz = sc*gy' + P*z + Q  */
int hwg_norm_bwd_reduce(const void* g, const void* z, const float* coef, float slope, int N, int H, int W, int C,
                        int kh, int kw, float* sums, void* stream);
int hwg_gn_bwd_coeffs(const float* sums, const float* save_mean_rstd, const float* gamma, int N, int C, int groups,
                      int64_t HW, float* spq, float* dgamma, float* dbeta, void* stream);
int hwg_norm_bwd_apply(const void* g, const void* z, const float* coef, const float* spq, float slope, int N, int H,
                       int W, int C, int kh, int kw, void* gz, void* stream);
/* SpectralNorm._update_u_v (:19-32) for all wrapped layers at once (three small launches over (layer, chunk)
 * grids): v = normalize(W^T u), u = normalize(W v) written back in place, inv_sigma[layer] = 1 / (u . W v).
 * jobs_dev: device array of { const float* w [h][wd]; float* u [h]; float* v [wd]; int32 h, wd } (32 bytes each);
 * max_h / max_wd: largest h / wd over the jobs; norms_scratch: 2*njobs floats, zeroed once by the caller (the last
 * launch leaves them zero).  The packed bf16 operands are then W * inv_sigma (hwgMapJob.scale_dev). */
int hwg_spectral_norm(const void* jobs_dev, int njobs, int max_h, int max_wd, float* norms_scratch,
                      float* inv_sigma, void* stream);
/* out[c] += sum over rows of x[row, c] (x bf16 [rows, C] NHWC): bias gradients of the discriminator's convolutions
 * (autograd's sum over (n, h, w) of the output gradient).  out is zeroed by the caller. */
int hwg_channel_sum(const void* x, int64_t rows, int C, float* out, void* stream);
/* Backward of SpectralNorm (:30-32: weight = w_bar / sigma, sigma = u . (W v), u and v constants) for all wrapped
 * layers (two launches): given gw = (dL/dweight) / sigma in the parameter layout (the unpacked wgrad, already scaled
 * by inv_sigma), in place  gw -= (<gw, w_bar> * inv_sigma[layer]) * u v^T.
 * jobs_dev: device array of { const float* w_bar; float* gw; const float* u; const float* v; int32 h, wd } (40 bytes);
 * max_elems = largest h*wd; dots_scratch: njobs floats, zeroed by the caller. */
int hwg_spectral_norm_bwd(const void* jobs_dev, int njobs, int64_t max_elems, const float* inv_sigma,
                          float* dots_scratch, void* stream);

/* ------------------------------------------------------------------------
 * Perceptual encoder (reference model/autoencoder.py:341-410 `Encoder2`; SURVEY.md 8 row f1, second half) and the
 * perceptual loss built on it (trainer/hw_with_style_trainer.py:740-748).  The convolutions, GroupNorm, AvgPool2d and
 * Dropout2d + ReLU passes are the discriminator's entry points above; these two are what they do not cover.
 * GPU parity: tests/test_enc_gpu.py (goldens of the unmodified reference).
 * ---------------------------------------------------------------------- */
/* The residual additions `x = self.conv1(x); x += res` (autoencoder.py:400-401, :404-405): y = a + b on NHWC bf16
 * [N,HW,C] (C in {16,32,64,128,256}; y may alias a or b) and, when stats != NULL (fp32 [N,C,2], zeroed by the caller),
 * the per-(n,c) sum and sum of squares of y that the GroupNorm which follows consumes (hwg_gn_coeffs). */
int hwg_add_stats(const void* a, const void* b, void* y, int N, int64_t HW, int C, float* stats, void* stream);
/* F.l1_loss between the halves of one feature tensor f = [orig ; recon] (trainer :743-747: torch.chunk(b, 2, dim=0)):
 * *loss += loss_scale * sum |recon - orig| (loss_scale = 1/half_numel for reduction='mean'), and, when g != NULL,
 * g[i] = grad_scale * sign(recon[i] - orig[i]) as bf16 (the gradient w.r.t. the recon half, sign(0) = 0 as in torch).
 * dtype: HWG_DT_BF16 or HWG_DT_F32 element type of f; half_numel % 8 == 0; 16-byte aligned pointers. */
int hwg_l1_halves(const void* f, int dtype, int64_t half_numel, float loss_scale, float grad_scale, float* loss,
                  void* g, void* stream);

/* ------------------------------------------------------------------------
 * DTW alignment of a label to the recognizer output — reference `correct_pred`, model/hw_with_style.py:18-74 (SURVEY.md 8
 * row f3; called by HWWithStyle.autoencode / extract_style, :279-291).  label_with_blanks = blank, c0, blank, c1, ...,
 * blank (L = 2S+1 columns); dtw[i][j] = (1 - pred[i-1, b, label_with_blanks[j-1]]) + min(dtw[i-1][j], dtw[i-1][j-1],
 * dtw[i][j-1]) inside the band |i - j| <= w, w = max(T/2, |T - L|) (first minimum wins, as torch.min); the path is
 * traced back from (T, L) and the label under it written front to back.
 * pred [T,B,C] fp32 contiguous; label int32, element (s,b) at label[s*label_stride_s + b*label_stride_b];
 * hist [B,T,L] bytes and scratch [B,T+L] int32 are workspaces; out [T+L,B] int32 (the caller zero-fills it: rows past
 * out_len[b] are the reference's zero padding), out_len [B].  One CTA per sequence, 2S+2 <= 1024.
 * Bit-exact against the reference's alignments: tests/test_dtw_gpu.py. */
int hwg_dtw_align(const float* pred, int T, int B, int C, const int32_t* label, int64_t label_stride_s,
                  int64_t label_stride_b, int S, uint8_t* hist, int32_t* out, int32_t* out_len, int32_t* scratch,
                  void* stream);

/* ------------------------------------------------------------------------
 * Text spacing (reference HWWithStyle.insert_spaces, model/hw_with_style.py:302-328):
 *   line b = for every character i < lengths[b]:  [blank] * round(N(counts[i,b,0], count_std))
 *                                               + [label[i,b]] * round(N(counts[i,b,1], dup_std))   (1 when n_out == 1)
 *   spaced [T,B,C] one-hot with T = max line length + max(ceil(max counts), 3), blank (class 0) behind every line.
 * z: the standard normals of the reference's np.random.normal calls, drawn by the host from the same stream in one call
 *    (order: line, character, count before duplicates); z_off[b] = index of line b's first draw.  The kernel evaluates
 *    loc + std * z in doubles with two roundings and rounds half to even — numpy's arithmetic and Python's round().
 * hwg_insert_spaces_plan: reps [B][L][2] = (blanks, repetitions), offsets [B][L+1] = exclusive prefix sums (entries from
 *    lengths[b] on hold the line length), info [B+1] = line lengths, then ceil(max over ALL entries of counts) (:303).
 * hwg_insert_spaces_fill: writes every element of spaced [T,B,C] fp32 (label: int32 or int64 [L,B] with element strides).
 * The caller reads info (ONE device->host copy of B+1 integers, instead of the reference's 2*L*B `.item()` calls) to size
 * `spaced`.  Integer work: bit-exact against the reference's goldens. */
int hwg_insert_spaces_plan(const int32_t* lengths, const float* counts, int n_out, const double* z, const int64_t* z_off,
                           int L, int B, double count_std, double dup_std, int32_t* reps, int32_t* offsets,
                           int32_t* info, void* stream);
int hwg_insert_spaces_fill(const void* label, int label_is_i64, int64_t label_stride_l, int64_t label_stride_b,
                           const int32_t* lengths, const int32_t* reps, const int32_t* offsets, int L, int B, int T,
                           int C, float* spaced, void* stream);

/* ------------------------------------------------------------------------
 * Peer-memory exchange for data-parallel BatchNorm (coupling 1).
 * The reference is single-process: nn.BatchNorm2d/1d in cnn_only_hwr.py:36,79
 * normalise with the statistics of the WHOLE batch.  With the batch sharded
 * over one process per GPU the per-channel sums are added over the ranks inside
 * the consuming kernel, through mailboxes in peer-mapped HBM (NVLink stores with
 * flags in the payload) — no NCCL launch, no stream fork/join, graph-capturable.
 *
 * The caller (dp.PeerExchange) owns all memory: a zero-filled mailbox of
 * hwg_peer_mailbox_bytes(world, slots) on every rank, mapped into every other
 * rank's address space (torch symmetric memory or CUDA IPC); peer_mailboxes is a
 * DEVICE array [world] of those base addresses as seen from this process;
 * epochs is a zero-filled device uint32 [slots]; fault a device int that a rank
 * sets (instead of hanging) when a peer has not answered within 10 s.
 * A slot identifies one call site; all ranks must use the same slot for the same
 * exchange and launch their exchanges in the same order. */
#define HWG_PEER_MAX_WORLD 16
#define HWG_PEER_SLOT_VALUES 1024 /* fp32 values one exchange can carry */
int64_t hwg_peer_mailbox_bytes(int world, int slots);
/* cudaDeviceEnablePeerAccess(current -> peer_device), tolerant of "already enabled" (CUDA-IPC mapping only). */
int hwg_peer_enable_access(int peer_device);
/* hwg_bn_coeffs (training mode) over the joint batch: folds the local [N,C,2] sums over n, adds them over the
 * ranks, then coefficients / saved (mean, rstd) / running statistics from global_count = elements per channel
 * over ALL ranks.  C <= 512.  One launch. */
int hwg_bn_coeffs_peer(const float* stats, int N, int C, int64_t global_count, const float* weight,
                       const float* bias, float* running_mean, float* running_var, float momentum, float eps,
                       float* coef, float* save_mean_rstd, const uint64_t* peer_mailboxes, int world, int rank,
                       int slot, int slots, uint32_t* epochs, int* fault, void* stream);
/* out[i] = sum over ranks of in[i], n <= HWG_PEER_SLOT_VALUES fp32 values, identical bits on every rank
 * (in == out allowed).  Used between hwg_bn_bwd_reduce and hwg_bn_bwd_apply. */
int hwg_peer_allreduce_f32(const float* in, float* out, int n, const uint64_t* peer_mailboxes, int world,
                           int rank, int slot, int slots, uint32_t* epochs, int* fault, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HWG_B200_H */
