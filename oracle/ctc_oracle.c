/*
 * ctc_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C, fp32) of the CTC loss the reference calls through
 * F.ctc_loss (reference model/loss.py:28-30) and of its best-path decode
 * (reference utils/string_utils.py:51-57).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object; the product path (libhwg_b200.so) never does.
 *
 * The arithmetic of F.ctc_loss is not under /root/reference: it is PyTorch
 * ATen (third-party, pinned here to torch 2.11.0), aten/src/ATen/native/
 * LossCTC.cpp: ctc_loss_cpu_template (alpha recursion, nll) and
 * ctc_loss_backward_cpu_template (beta recursion, per-class log-sum-exp of
 * alpha+beta, grad = (exp(lp) - exp(lcab + nll - lp)) * grad_nll).  This file
 * restates that published algorithm; tests/golden/ctc_*.npz (made by
 * oracle/make_golden.py from torch's own CPU kernel on the reference's call
 * pattern) pin it.
 *
 * Layouts: log_probs [T,B,C]; targets element (b,s) at tgt[b*sb + s*ss];
 * log_alpha/log_beta [B,T,L], L = 2*S_max+1; grad [T,B,C].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline int aug(const int32_t* tgt, int64_t sb, int64_t ss, int b, int s, int blank) {
  return (s & 1) ? tgt[b * sb + (int64_t)(s >> 1) * ss] : blank;
}

static inline float lse3f(float a, float b, float c) {
  float m = a > b ? a : b;
  m = m > c ? m : c;
  if (m == -INFINITY) m = 0.f; /* LossCTC.cpp: "if (lamax == neginf) lamax = 0" */
  return logf(expf(a - m) + expf(b - m) + expf(c - m)) + m;
}

/* LossCTC.cpp ctc_loss_cpu_template */
void ctc_oracle_forward(const float* lp, int T, int B, int C, const int32_t* tgt, int64_t sb,
                        int64_t ss, int S_max, const int32_t* in_len, const int32_t* tg_len,
                        int blank, float* nll, float* log_alpha) {
  const int L = 2 * S_max + 1;
  for (int b = 0; b < B; ++b) {
    const int Tb = in_len[b], Sb = tg_len[b], Lb = 2 * Sb + 1;
    float* A = log_alpha + (size_t)b * T * L;
    for (size_t i = 0; i < (size_t)T * L; ++i) A[i] = -INFINITY;
    if (Tb <= 0) { nll[b] = INFINITY; continue; }
    A[0] = lp[((size_t)0 * B + b) * C + blank];
    if (Sb > 0) A[1] = lp[((size_t)0 * B + b) * C + aug(tgt, sb, ss, b, 1, blank)];
    for (int t = 1; t < Tb; ++t) {
      const float* row = lp + ((size_t)t * B + b) * C;
      const float* P = A + (size_t)(t - 1) * L;
      float* Q = A + (size_t)t * L;
      for (int s = 0; s < Lb; ++s) {
        int l = aug(tgt, sb, ss, b, s, blank);
        float a1 = P[s];
        float a2 = s > 0 ? P[s - 1] : -INFINITY;
        float a3 = (s > 1 && aug(tgt, sb, ss, b, s - 2, blank) != l) ? P[s - 2] : -INFINITY;
        Q[s] = lse3f(a1, a2, a3) + row[l];
      }
    }
    const float* last = A + (size_t)(Tb - 1) * L;
    float l1 = last[Lb - 1];
    float l2 = Sb > 0 ? last[Lb - 2] : -INFINITY;
    float m = l1 > l2 ? l1 : l2;
    if (m == -INFINITY) m = 0.f;
    nll[b] = -(logf(expf(l1 - m) + expf(l2 - m)) + m);
  }
}

/* reduction='mean' (LossCTC.cpp ctc_loss: mean(nll / clamp_min(target_lengths,1)))
 * followed by the reference wrapper's inf -> 0 (model/loss.py:30). */
float ctc_oracle_reduce_mean(const float* nll, const int32_t* tg_len, int B) {
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += nll[b] / (float)(tg_len[b] < 1 ? 1 : tg_len[b]);
  acc /= (float)B;
  return isinf(acc) ? 0.f : acc;
}

/* LossCTC.cpp ctc_loss_backward_cpu_template; grad_nll[b] is the gradient that
 * reaches nll[b] (= grad_out/(B*max(S_b,1)) for reduction='mean'). */
void ctc_oracle_backward(const float* grad_nll, const float* lp, int T, int B, int C,
                         const int32_t* tgt, int64_t sb, int64_t ss, int S_max,
                         const int32_t* in_len, const int32_t* tg_len, int blank,
                         const float* nll, const float* log_alpha, float* log_beta,
                         float* grad) {
  const int L = 2 * S_max + 1;
  for (size_t i = 0; i < (size_t)T * B * C; ++i) grad[i] = -INFINITY;
  for (int b = 0; b < B; ++b) {
    const int Tb = in_len[b], Sb = tg_len[b], Lb = 2 * Sb + 1;
    const float* A = log_alpha + (size_t)b * T * L;
    float* Bt = log_beta + (size_t)b * T * L;
    for (size_t i = 0; i < (size_t)T * L; ++i) Bt[i] = -INFINITY;
    if (Tb > 0) {
      int t = Tb - 1;
      const float* row = lp + ((size_t)t * B + b) * C;
      float* g = grad + ((size_t)t * B + b) * C;
      Bt[(size_t)t * L + Lb - 1] = row[blank];
      g[blank] = A[(size_t)t * L + Lb - 1] + Bt[(size_t)t * L + Lb - 1];
      if (Sb > 0) {
        int l = aug(tgt, sb, ss, b, Lb - 2, blank);
        Bt[(size_t)t * L + Lb - 2] = row[l];
        g[l] = A[(size_t)t * L + Lb - 2] + Bt[(size_t)t * L + Lb - 2];
      }
    }
    for (int t = Tb - 2; t >= 0; --t) {
      const float* row = lp + ((size_t)t * B + b) * C;
      float* g = grad + ((size_t)t * B + b) * C;
      const float* N = Bt + (size_t)(t + 1) * L;
      float* Q = Bt + (size_t)t * L;
      for (int s = 0; s < Lb; ++s) {
        int l = aug(tgt, sb, ss, b, s, blank);
        float b1 = N[s];
        float b2 = s < Lb - 1 ? N[s + 1] : -INFINITY;
        float b3 = (s < Lb - 2 && aug(tgt, sb, ss, b, s + 2, blank) != l) ? N[s + 2] : -INFINITY;
        Q[s] = lse3f(b1, b2, b3) + row[l];
        float ab = A[(size_t)t * L + s] + Q[s];
        float* lcab = &g[l];
        if (*lcab == -INFINITY) {
          *lcab = ab;
        } else {
          float m = *lcab > ab ? *lcab : ab;
          *lcab = logf(expf(*lcab - m) + expf(ab - m)) + m;
        }
      }
    }
    for (int t = 0; t < T; ++t) {
      const float* row = lp + ((size_t)t * B + b) * C;
      float* g = grad + ((size_t)t * B + b) * C;
      if (t < Tb)
        for (int c = 0; c < C; ++c) g[c] = (expf(row[c]) - expf(g[c] + nll[b] - row[c])) * grad_nll[b];
      else
        for (int c = 0; c < C; ++c) g[c] = 0.f;
    }
  }
}

/* utils/string_utils.py:51-57 naive_decode, per sequence b of a [T,B,C] array. */
void ctc_oracle_decode(const float* lp, int T, int B, int C, const int32_t* in_len, int blank,
                       int32_t* raw, int32_t* decoded, int32_t* decoded_len) {
  for (int b = 0; b < B; ++b) {
    int n = 0, Tb = in_len ? in_len[b] : T;
    for (int t = 0; t < T; ++t) {
      const float* row = lp + ((size_t)t * B + b) * C;
      int bi = 0;
      for (int c = 1; c < C; ++c) if (row[c] > row[bi]) bi = c; /* first maximum wins */
      raw[(size_t)t * B + b] = bi;
      if (t < Tb && bi != blank && !(t > 0 && bi == raw[(size_t)(t - 1) * B + b]))
        decoded[(size_t)b * T + n++] = bi;
    }
    decoded_len[b] = n;
  }
}
