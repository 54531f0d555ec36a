// Fused optimizer tail over flat parameter buffers: global grad-norm, then ONE pass doing gradient clipping,
// AdamW, the EMA update and the refresh of the bf16 shadow weights the GEMMs read.
// 36 B/param algorithmic (read p,g,m,v,ema; write p,m,v,ema) + 2 B/param shadow + 4 B/param norm pass.
//
// Reference: /root/reference/image/train.py:94-105 (update_ema), 253-259 (AdamW), 402-412 (clip, step, EMA);
// torch.nn.utils.clip_grad_norm_ (coef = max_norm / (norm + 1e-6), clamped to 1) and torch.optim.AdamW math.
#include <stdlib.h>
#include "common.cuh"

namespace reed {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n4, int64_t n,
                                                     double* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 v = load4(g + i * 4);
    acc += (v.v[0] * v.v[0] + v.v[1] * v.v[1]) + (v.v[2] * v.v[2] + v.v[3] * v.v[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 * 4; i < n; ++i) acc += g[i] * g[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, (double)v);
  }
}

struct AdamArgs {
  float* p; const float* g; float* m; float* v; float* ema; bf16* shadow;
  int64_t n;
  const double* norm_sq;   // device scalar: sum of squares of ALL gradients (after all-reduce), or null = no clipping
  const int* step_dev;     // device scalar holding the 1-based step (CUDA-graph replays read it), or null = host `step`
  float max_norm, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt, ema_decay, grad_scale;
};

__device__ __forceinline__ void adam_one(const AdamArgs& a, float coef, float& p, float g, float& m, float& v,
                                         float& e) {
  g *= coef;
  p *= (1.f - a.lr * a.weight_decay);
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
  p -= (a.lr / a.bias_c1) * (m / denom);
  e = a.ema_decay * e + (1.f - a.ema_decay) * p;
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(AdamArgs a) {
  if (a.step_dev != nullptr) {     // same double-precision bias corrections the host computes for the scalar-step call
    const double st = (double)*a.step_dev;
    a.bias_c1 = (float)(1.0 - pow((double)a.beta1, st));
    a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, st));
  }
  float coef = a.grad_scale;
  if (a.norm_sq != nullptr) {
    float norm = (float)sqrt(*a.norm_sq) * a.grad_scale;
    coef *= fminf(a.max_norm / (norm + 1e-6f), 1.f);
  }
  const int64_t n4 = a.n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 p = load4(a.p + i * 4), g = load4(a.g + i * 4), m = load4(a.m + i * 4), v = load4(a.v + i * 4),
       e = load4(a.ema + i * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) adam_one(a, coef, p.v[j], g.v[j], m.v[j], v.v[j], e.v[j]);
    store4(a.p + i * 4, p);
    store4(a.m + i * 4, m);
    store4(a.v + i * 4, v);
    store4(a.ema + i * 4, e);
    if (a.shadow) store4(a.shadow + i * 4, p);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 * 4; i < a.n; ++i) {
      float p = a.p[i], m = a.m[i], v = a.v[i], e = a.ema[i];
      adam_one(a, coef, p, a.g[i], m, v, e);
      a.p[i] = p; a.m[i] = m; a.v[i] = v; a.ema[i] = e;
      if (a.shadow) a.shadow[i] = __float2bfloat16_rn(p);
    }
  }
}

__global__ void __launch_bounds__(256) ema_kernel(const float* __restrict__ p, float* __restrict__ ema, int64_t n,
                                                   float decay) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    ema[i] = decay * ema[i] + (1.f - decay) * p[i];
}

}  // namespace reed

using namespace reed;

static inline int opt_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  return (int)(b < 1 ? 1 : (b > kNumSMs * 8 ? kNumSMs * 8 : b));
}

// out (double, device) += sum g^2 ; caller zeroes `out` once per step
extern "C" int reed_grad_sumsq(const void* g, int64_t n, void* out, void* stream) {
  if (n == 0) return 0;
  REED_REQUIRE(((uintptr_t)g & 15) == 0, "grad_sumsq: buffer must be 16-byte aligned");
  sumsq_kernel<<<opt_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float*)g, n / 4, n, (double*)out);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_adamw_ema(void* p, const void* g, void* m, void* v, void* ema, void* shadow_bf16, int64_t n,
                              const void* norm_sq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                              float eps, float weight_decay, int step, float ema_decay, const void* step_dev,
                              void* stream) {
  if (n == 0) return 0;
  REED_REQUIRE(step >= 1 || step_dev != nullptr, "adamw: step is 1-based");
  REED_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) & 15) == 0 &&
                   ((uintptr_t)shadow_bf16 & 7) == 0,
               "adamw: buffers must be 16-byte aligned");
  AdamArgs a;
  a.p = (float*)p; a.g = (const float*)g; a.m = (float*)m; a.v = (float*)v; a.ema = (float*)ema;
  a.shadow = (bf16*)shadow_bf16; a.n = n; a.norm_sq = (const double*)norm_sq;
  a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.ema_decay = ema_decay; a.grad_scale = grad_scale; a.step_dev = (const int*)step_dev;
  adamw_ema_kernel<<<opt_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(a);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_ema_update(const void* p, void* ema, int64_t n, float decay, void* stream) {
  if (n == 0) return 0;
  int64_t b = (n + 255) / 256;
  ema_kernel<<<(int)(b > kNumSMs * 8 ? kNumSMs * 8 : b), 256, 0, (cudaStream_t)stream>>>((const float*)p, (float*)ema, n, decay);
  REED_LAUNCH_CHECK();
  return 0;
}
