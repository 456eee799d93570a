// Flat fused Adam step (trainer/hw_with_style_trainer.py:381-391 + base_trainer.py:95-100):
// clip_grad_value_ + torch.optim.Adam(lr, betas, eps, weight_decay=0) over ONE flat fp32 buffer holding every
// parameter of an optimizer group, in one launch, with the step counter on the device (CUDA-graph replayable) and the
// gradient zeroed on the way out (replaces ~10 multi-tensor launches + zero_grad per step).
#include "common.cuh"

namespace hwg {

__global__ void adam_count_kernel(float* step) { step[0] += 1.f; }

__global__ void __launch_bounds__(256) adam_flat_kernel(float4* __restrict__ p, float4* __restrict__ g,
                                                        float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                        float lr, float b1, float b2, float eps, float clip,
                                                        float grad_scale, const float* __restrict__ step, int zero_grad) {
  const float t = step[0];
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 P = p[i], G = g[i], M = m[i], V = v[i];
    float* pp = reinterpret_cast<float*>(&P);
    float* gg = reinterpret_cast<float*>(&G);
    float* mm = reinterpret_cast<float*>(&M);
    float* vv = reinterpret_cast<float*>(&V);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = gg[k] * grad_scale;
      if (clip > 0.f) gr = fminf(fmaxf(gr, -clip), clip);
      mm[k] = mm[k] + (gr - mm[k]) * (1.f - b1);                 // exp_avg.lerp_(grad, 1 - beta1)
      vv[k] = vv[k] * b2 + gr * gr * (1.f - b2);                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vv[k]) * inv_sqrt_bc2 + eps;
      pp[k] -= step_size * (mm[k] / denom);
    }
    p[i] = P; m[i] = M; v[i] = V;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace hwg

extern "C" int hwg_adam_flat(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                             float eps, float clip_value, float grad_scale, float* step_dev, int zero_grad,
                             void* stream) {
  HWG_REQUIRE(p && g && m && v && step_dev && n > 0 && n % 4 == 0, "hwg_adam_flat: bad argument (n must be a multiple of 4)");
  HWG_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "hwg_adam_flat: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  hwg::adam_count_kernel<<<1, 1, 0, s>>>(step_dev);
  int rc = hwg::check_launch("adam_count_kernel");
  if (rc) return rc;
  const long long n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  hwg::adam_flat_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g),
                                               reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4, lr, beta1,
                                               beta2, eps, clip_value, grad_scale, step_dev, zero_grad);
  return hwg::check_launch("adam_flat_kernel");
}

// ---- backward of the small fp32 dense layers of the style path (style MLP pure_gen.py:31-39, AdaIN projections
// pure_gen.py:57,63): y = act(x W^T + b).  One launch: blocks [0,O) produce row o of g_W and g_b[o] (threads over k,
// loop over the batch); blocks [O, O+B) produce row b of g_x: LB_PARTS groups of threads each reduce a slice of the
// O outputs for every k (W read along k, coalesced), combined through shared memory.
namespace hwg {
constexpr int LB_THREADS = 512, LB_OSLICE = 128;

__global__ void __launch_bounds__(LB_THREADS) linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                const float* __restrict__ gy, const float* __restrict__ W,
                                                                int B, int K, int O, int act, float slope,
                                                                float* __restrict__ gx, float* __restrict__ gW,
                                                                float* __restrict__ gb, int accumulate) {
  extern __shared__ float sm[];   // g_pre of this block's row: [max(B,O)], then the partial sums [parts][K]
  const int blk = blockIdx.x;
  if (blk < O) {
    const int o = blk;
    float* gp = sm;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      float g = gy[(size_t)b * O + o];
      if (act == HWG_ACT_LRELU) g *= (y[(size_t)b * O + o] > 0.f) ? 1.f : slope;
      gp[b] = g;
    }
    __syncthreads();
    if (gW) {
      // threads = (k, slice of the batch): K = 128 would leave 3/4 of the block idle on one long chain of dependent
      // L2-latency loads (32 us per launch at B = 128, 21 launches per step); the slices meet in shared memory
      const int kw = K < (int)blockDim.x ? K : (int)blockDim.x, parts = blockDim.x / kw;
      const int k0 = threadIdx.x % kw, pr = threadIdx.x / kw;
      float* part = sm + (B > O ? B : O);
      for (int kb = 0; kb < K; kb += kw) {
        const int k = kb + k0;
        float acc = 0.f;
        if (pr < parts && k < K) {
#pragma unroll 8
          for (int b = pr; b < B; b += parts) acc = fmaf(gp[b], x[(size_t)b * K + k], acc);
        }
        if (pr < parts) part[pr * kw + k0] = acc;
        __syncthreads();
        if (pr == 0 && k < K) {
          float t = 0.f;
          for (int q = 0; q < parts; ++q) t += part[q * kw + k0];
          float* d = gW + (size_t)o * K + k;
          *d = accumulate ? *d + t : t;
        }
        __syncthreads();
      }
    }
    if (gb && threadIdx.x == 0) {
      float acc = 0.f;
      for (int b = 0; b < B; ++b) acc += gp[b];
      gb[o] = accumulate ? gb[o] + acc : acc;
    }
  } else if (gx) {
    // block = (batch row b, slice of LB_OSLICE outputs): partial g_x[b, :] over the slice, added atomically
    // (g_x is zeroed by the caller) — O = 1984 for the concatenated AdaIN projections would otherwise be one
    // long serial loop on a handful of blocks
    const int nsl = (O + LB_OSLICE - 1) / LB_OSLICE;
    const int b = (blk - O) / nsl, o0 = ((blk - O) % nsl) * LB_OSLICE;
    const int on = min(LB_OSLICE, O - o0);
    float* gp = sm;
    const int nmax = B > O ? B : O;
    float* part = sm + nmax;
    for (int o = threadIdx.x; o < on; o += blockDim.x) {
      float g = gy[(size_t)b * O + o0 + o];
      if (act == HWG_ACT_LRELU) g *= (y[(size_t)b * O + o0 + o] > 0.f) ? 1.f : slope;
      gp[o] = g;
    }
    __syncthreads();
    const int kw = K < (int)blockDim.x ? K : (int)blockDim.x;   // threads along k
    const int parts = blockDim.x / kw;                            // groups along o
    const int k0 = threadIdx.x % kw, pr = threadIdx.x / kw;
    for (int kb = 0; kb < K; kb += kw) {
      const int k = kb + k0;
      float acc = 0.f;
      if (pr < parts && k < K) {
#pragma unroll 8
        for (int o = pr; o < on; o += parts) acc = fmaf(gp[o], W[(size_t)(o0 + o) * K + k], acc);
      }
      if (pr < parts) part[pr * kw + k0] = acc;
      __syncthreads();
      if (pr == 0 && k < K) {
        float t = 0.f;
        for (int q = 0; q < parts; ++q) t += part[q * kw + k0];
        atomicAdd(&gx[(size_t)b * K + k], t);
      }
      __syncthreads();
    }
  }
}
}  // namespace hwg

extern "C" int hwg_linear_bwd_f32(const float* x, const float* y, const float* gy, const float* W, int B, int K,
                                  int O, int act, float slope, float* gx, float* gW, float* gb, int accumulate,
                                  void* stream) {
  HWG_REQUIRE(x && gy && W && B > 0 && K > 0 && O > 0, "hwg_linear_bwd_f32: bad argument");
  HWG_REQUIRE(act == 0 || (act == HWG_ACT_LRELU && y), "hwg_linear_bwd_f32: activation must be none or LeakyReLU (needs y)");
  const int nmax = B > O ? B : O;
  HWG_REQUIRE(nmax <= 8192, "hwg_linear_bwd_f32: B and O must be <= 8192");
  const size_t smem = (size_t)(nmax + hwg::LB_THREADS) * sizeof(float);
  const int nsl = (O + hwg::LB_OSLICE - 1) / hwg::LB_OSLICE;
  hwg::linear_bwd_kernel<<<O + (gx ? B * nsl : 0), hwg::LB_THREADS, smem, (cudaStream_t)stream>>>(
      x, y, gy, W, B, K, O, act, slope, gx, gW, gb, accumulate);
  return hwg::check_launch("linear_bwd_kernel");
}
