// Gradient exchange of the data-parallel step over NVLink 5 / NVSwitch multicast (NVLS), fused with the arithmetic on
// either side of it.  Two kernels, both working on buffers registered as symmetric memory with a multicast mapping
// (torch.distributed._symmetric_memory; the host side is reed_b200/image/nvls.py):
//
//   reed_nvls_reduce_scatter_sumsq   gradient reduce-scatter + global-norm partial in one pass: every rank pulls the
//       SUM over ranks of the slice it owns with multimem.ld_reduce (the switch adds the peers' copies in flight - no
//       staging buffers, no reduction on the SMs), stores it over its local copy of the slice and accumulates the
//       slice's sum of squares for clip_grad_norm_.
//   reed_adamw_ema_mc                the fused clip + AdamW + EMA pass of optim.cu on the owned slice, whose bf16 GEMM
//       operands are written with multimem.st through the multicast address: one store lands in every rank's shadow
//       buffer, i.e. the all-gather of the updated weights rides in the optimizer kernel's own stores.
//
// Reference semantics: DistributedDataParallel's gradient all-reduce(average) set up by accelerate
// (/root/reference/image/train.py:151,293,401) followed by clip_grad_norm_ / AdamW / update_ema (train.py:94-105,
// 253-259,402-412).  Sum over ranks here, the 1/world factor is folded into the optimizer kernel (grad_scale).
//
// Ordering is the caller's job (nvls.py): a cross-rank barrier on the stream before the reduce-scatter (all ranks have
// finished writing the bucket's gradients) and before the next forward (all ranks' multicast operand stores are done).
// Validated on 2 x B200 in round 2 (profiles/r02_sharded_check.txt); the trainer's default exchange for world > 1.
#include <math.h>
#include "common.cuh"

namespace reed {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
// 4 bf16 (8 bytes) to the same offset of every rank's buffer
__device__ __forceinline__ void multimem_st_bf16x4(bf16* mc, const F4& f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f.v[0], f.v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f.v[2], f.v[3]);
  asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};"
               :
               : "l"(mc), "r"(*reinterpret_cast<uint32_t*>(&a)), "r"(*reinterpret_cast<uint32_t*>(&b))
               : "memory");
}

constexpr int kRsUnroll = 4;   // multimem loads in flight per thread: the round trip through the switch is microseconds

// local[lo .. lo+n) = sum over ranks of grad[lo .. lo+n);  *norm_sq += sum of squares of the result.  n % 4 == 0.
__global__ void __launch_bounds__(512) nvls_reduce_scatter_sumsq_kernel(const float* __restrict__ mc, float* __restrict__ local,
                                                                         int64_t lo, int64_t n4, double* __restrict__ norm_sq) {
  __shared__ float red[16];
  const float* src = mc + lo;
  float* dst = local + lo;
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kRsUnroll - 1) * stride < n4; i += kRsUnroll * stride) {
    float4 v[kRsUnroll];
#pragma unroll
    for (int u = 0; u < kRsUnroll; ++u) v[u] = multimem_ld_reduce_add(src + (i + u * stride) * 4);
#pragma unroll
    for (int u = 0; u < kRsUnroll; ++u) {
      *reinterpret_cast<float4*>(dst + (i + u * stride) * 4) = v[u];
      acc += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
    }
  }
  for (; i < n4; i += stride) {
    float4 v = multimem_ld_reduce_add(src + i * 4);
    *reinterpret_cast<float4*>(dst + i * 4) = v;
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && norm_sq != nullptr) atomicAdd(norm_sq, (double)v);
  }
}

struct AdamMcArgs {
  float* p; const float* g; float* m; float* v; float* ema;
  bf16* shadow_mc;         // multicast address of the bf16 operand buffer, already offset to this rank's slice
  int64_t n;
  const double* norm_sq;   // device scalar: global sum of squares (after the cross-rank sum), or null = no clipping
  const int* step_dev;     // device scalar holding the 1-based step, or null = host `step`
  float max_norm, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt, ema_decay, grad_scale;
};

// Same arithmetic, in the same order, as adamw_ema_kernel / adam_one of optim.cu.
__global__ void __launch_bounds__(256) adamw_ema_mc_kernel(AdamMcArgs a) {
  if (a.step_dev != nullptr) {
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
    for (int j = 0; j < 4; ++j) {
      float gj = g.v[j] * coef;
      p.v[j] *= (1.f - a.lr * a.weight_decay);
      m.v[j] = a.beta1 * m.v[j] + (1.f - a.beta1) * gj;
      v.v[j] = a.beta2 * v.v[j] + (1.f - a.beta2) * gj * gj;
      float denom = sqrtf(v.v[j]) / a.bias_c2_sqrt + a.eps;
      p.v[j] -= (a.lr / a.bias_c1) * (m.v[j] / denom);
      e.v[j] = a.ema_decay * e.v[j] + (1.f - a.ema_decay) * p.v[j];
    }
    store4(a.p + i * 4, p);
    store4(a.m + i * 4, m);
    store4(a.v + i * 4, v);
    store4(a.ema + i * 4, e);
    multimem_st_bf16x4(a.shadow_mc + i * 4, p);
  }
}

}  // namespace reed

using namespace reed;

extern "C" int reed_nvls_reduce_scatter_sumsq(const void* grad_multicast, void* grad_local, int64_t lo, int64_t n,
                                              void* norm_sq, int ctas, void* stream) {
  if (n == 0) return 0;
  REED_REQUIRE(grad_multicast != nullptr && grad_local != nullptr, "nvls_reduce_scatter: buffers missing");
  REED_REQUIRE(lo >= 0 && lo % 4 == 0 && n % 4 == 0, "nvls_reduce_scatter: slice must be 16-byte aligned (lo=%lld n=%lld)",
               (long long)lo, (long long)n);
  REED_REQUIRE((((uintptr_t)grad_multicast | (uintptr_t)grad_local) & 15) == 0, "nvls_reduce_scatter: buffers must be 16-byte aligned");
  REED_REQUIRE(ctas >= 1 && ctas <= kNumSMs, "nvls_reduce_scatter: ctas %d out of range", ctas);
  nvls_reduce_scatter_sumsq_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>((const float*)grad_multicast, (float*)grad_local,
                                                                           lo, n / 4, (double*)norm_sq);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_adamw_ema_mc(void* p, const void* g, void* m, void* v, void* ema, void* shadow_multicast, int64_t n,
                                 const void* norm_sq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int step, float ema_decay, const void* step_dev,
                                 void* stream) {
  if (n == 0) return 0;
  REED_REQUIRE(step >= 1 || step_dev != nullptr, "adamw_mc: step is 1-based");
  REED_REQUIRE(n % 4 == 0 && shadow_multicast != nullptr, "adamw_mc: slice must be a multiple of 4 elements with a multicast operand buffer");
  REED_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) & 15) == 0 &&
                   ((uintptr_t)shadow_multicast & 7) == 0,
               "adamw_mc: buffers must be 16-byte aligned");
  AdamMcArgs a;
  a.p = (float*)p; a.g = (const float*)g; a.m = (float*)m; a.v = (float*)v; a.ema = (float*)ema;
  a.shadow_mc = (bf16*)shadow_multicast; a.n = n; a.norm_sq = (const double*)norm_sq;
  a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.ema_decay = ema_decay; a.grad_scale = grad_scale; a.step_dev = (const int*)step_dev;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  adamw_ema_mc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  REED_LAUNCH_CHECK();
  return 0;
}
