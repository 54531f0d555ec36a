// Shared device/host helpers for libreed_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace reed {

typedef __nv_bfloat16 bf16;

enum DType { kF32 = 0, kBF16 = 1 };

// thread-local last-error string surfaced through reed_last_error()
extern thread_local char g_err[512];
int fail(const char* fmt, ...);

#define REED_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return ::reed::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr,    \
                                               cudaGetErrorString(_e));                         \
  } while (0)

#define REED_REQUIRE(cond, ...)                                                                 \
  do {                                                                                          \
    if (!(cond)) return ::reed::fail(__VA_ARGS__);                                              \
  } while (0)

#define REED_LAUNCH_CHECK() REED_CHECK_CUDA(cudaGetLastError())

constexpr int kNumSMs = 148;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 4-wide vector access (16 B for float, 8 B for bf16); pointers must be suitably aligned ----
struct F4 { float v[4]; };

__device__ __forceinline__ F4 load4(const float* p) {
  float4 t = *reinterpret_cast<const float4*>(p);
  return F4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ F4 load4(const bf16* p) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  return F4{{__low2float(a), __high2float(a), __low2float(b), __high2float(b)}};
}
__device__ __forceinline__ void store4(float* p, const F4& f) {
  *reinterpret_cast<float4*>(p) = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const F4& f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f.v[0], f.v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f.v[2], f.v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- activations (match torch: GELU tanh approximation, SiLU) ----
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float u = k0 * (x + k1 * x * x2);
  float th = tanhf(u);
  float du = k0 * (1.f + 3.f * k1 * x2);
  return 0.5f * (1.f + th) + 0.5f * x * (1.f - th * th) * du;
}
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) {
  float s = 1.f / (1.f + expf(-x));
  return s * (1.f + x * (1.f - s));
}

// ---- GEMM epilogue selector shared by the SIMT and tcgen05 kernels ----
enum Epilogue {
  kEpiNone = 0,     // D = acc (+bias)            [optionally D += ... when accumulate]
  kEpiGelu = 1,     // h = acc+bias -> out2 ; D = gelu_tanh(h)
  kEpiSilu = 2,     // h = acc+bias -> out2 ; D = silu(h)
  kEpiGateRes = 3,  // y = acc+bias -> out2 ; D = res + gate[row/rows_per_group] * y   (D, res fp32)
  kEpiDGelu = 4,    // D = acc * gelu_tanh'(aux[row, col])
  kEpiDSilu = 5,    // D = acc * silu'(aux[row, col])
};

struct EpiParams {
  int kind;
  const float* bias;      // [N] or null
  const void* aux;        // kEpiGateRes: res fp32 [M, ld_aux]; kEpiDGelu/DSilu: pre-activation (act dtype) [M, ld_aux]
  int64_t ld_aux;
  const float* gate;      // kEpiGateRes: [groups, ld_gate] fp32
  int64_t ld_gate;
  int rows_per_group;
  void* out2;             // act dtype [M, ld_out2] or null
  int64_t ld_out2;
  int accumulate;         // kEpiNone with fp32 D only
  // Weight gradient that carries its bias gradient (tcgen05 path, kEpiNone, fp32 D): the B operand has a column of
  // ones at index n_store, so column n_store of the product is sum_k A[k, row] - it is added to bias_grad[row]
  // (atomically; the caller zeroes it) and columns >= n_store are not stored to D.  null = ordinary GEMM.
  float* bias_grad;
  int n_store;
};

// Applies the epilogue to 4 consecutive columns of one row.  TD = type of D, TA = activation type.
template <typename TD, typename TA>
__device__ __forceinline__ void epilogue_store4(const EpiParams& ep, TD* __restrict__ D, int64_t ldd, int row,
                                                int col, F4 acc) {
  if (ep.bias != nullptr) {
    F4 b = load4(ep.bias + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc.v[i] += b.v[i];
  }
  TD* dptr = D + (int64_t)row * ldd + col;
  switch (ep.kind) {
    case kEpiNone: {
      if (ep.accumulate) {
        F4 old = load4(dptr);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc.v[i] += old.v[i];
      }
      store4(dptr, acc);
    } break;
    case kEpiGelu:
    case kEpiSilu: {
      if (ep.out2) store4(reinterpret_cast<TA*>(ep.out2) + (int64_t)row * ep.ld_out2 + col, acc);
      F4 o;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // the saved pre-activation is rounded to TA; activate the rounded value so bwd sees the same h
        float h = to_f(from_f<TA>(acc.v[i]));
        o.v[i] = ep.kind == kEpiGelu ? gelu_tanh(h) : silu(h);
      }
      store4(dptr, o);
    } break;
    case kEpiGateRes: {
      if (ep.out2) store4(reinterpret_cast<TA*>(ep.out2) + (int64_t)row * ep.ld_out2 + col, acc);
      F4 r = load4(reinterpret_cast<const float*>(ep.aux) + (int64_t)row * ep.ld_aux + col);
      F4 g = load4(ep.gate + (int64_t)(row / ep.rows_per_group) * ep.ld_gate + col);
      F4 o;
#pragma unroll
      for (int i = 0; i < 4; ++i) o.v[i] = r.v[i] + g.v[i] * to_f(from_f<TA>(acc.v[i]));
      store4(dptr, o);
    } break;
    case kEpiDGelu:
    case kEpiDSilu: {
      F4 h = load4(reinterpret_cast<const TA*>(ep.aux) + (int64_t)row * ep.ld_aux + col);
      F4 o;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o.v[i] = acc.v[i] * (ep.kind == kEpiDGelu ? gelu_tanh_grad(h.v[i]) : silu_grad(h.v[i]));
      store4(dptr, o);
    } break;
  }
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (sm_90+).  pdl_launch(): this grid no longer holds back a dependent grid that was
// launched with the programmatic-serialization attribute - its CTAs may become resident (and run their prologue) as
// soon as SMs free up.  pdl_wait(): returns once every grid this one depends on has completed and its writes are
// visible; a no-op for a grid launched without the attribute.  A kernel launched WITH the attribute must call
// pdl_wait() before its first global-memory access.
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace reed
