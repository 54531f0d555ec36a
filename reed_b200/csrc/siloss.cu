// Fused SILoss kernels (north_star kernel 4): linear/cosine interpolant, velocity-target MSE and the
// per-token negative-cosine alignment, forward and backward.  HBM-bound: every tensor is read once per
// pass with 8/16-byte vector accesses and reduced with warp shuffles.
//
// Reference: /root/reference/image/loss.py:49-64 (interpolant), 175-186 (x_t, target, per-sample MSE),
// 204-225 (F.normalize both sides, eps 1e-12; -(z . z~).sum(-1).mean(-1)).
#include "common.cuh"

namespace reed {

__device__ __forceinline__ void path_coeffs(int path, float t, float& a, float& s, float& da, float& ds) {
  if (path == 0) {            // linear
    a = 1.f - t; s = t; da = -1.f; ds = 1.f;
  } else {                    // cosine
    const float hp = 1.5707963267948966f;
    float sn, cs;
    sincosf(t * hp, &sn, &cs);
    a = cs; s = sn; da = -hp * sn; ds = hp * cs;
  }
}

// x_t = alpha x + sigma eps   (per-sample coefficients);  grid.y = sample
__global__ void __launch_bounds__(256) interp_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                      const float* __restrict__ t, float* __restrict__ xt,
                                                      int per_sample, int path) {
  const int b = blockIdx.y;
  float a, s, da, ds;
  path_coeffs(path, t[b], a, s, da, ds);
  const int64_t base = (int64_t)b * per_sample;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < per_sample; i += gridDim.x * blockDim.x * 4) {
    F4 xv = load4(x + base + i), ev = load4(eps + base + i), o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = a * xv.v[j] + s * ev.v[j];
    store4(xt + base + i, o);
  }
}

// denoise[b] = mean((pred - (dalpha x + dsigma eps))^2); one CTA per sample
__global__ void __launch_bounds__(256) mse_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ x,
                                                       const float* __restrict__ eps, const float* __restrict__ t,
                                                       float* __restrict__ denoise, int per_sample, int path) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  float a, s, da, ds;
  path_coeffs(path, t[b], a, s, da, ds);
  const int64_t base = (int64_t)b * per_sample;
  float acc = 0.f;
  for (int i = threadIdx.x * 4; i < per_sample; i += blockDim.x * 4) {
    F4 p = load4(pred + base + i), xv = load4(x + base + i), ev = load4(eps + base + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float d = p.v[j] - (da * xv.v[j] + ds * ev.v[j]);
      acc += d * d;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) denoise[b] = v / per_sample;
  }
}

// dpred = g[b] * 2 (pred - target) / per_sample
__global__ void __launch_bounds__(256) mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ x,
                                                       const float* __restrict__ eps, const float* __restrict__ t,
                                                       const float* __restrict__ g, float* __restrict__ dpred,
                                                       int per_sample, int path) {
  const int b = blockIdx.y;
  float a, s, da, ds;
  path_coeffs(path, t[b], a, s, da, ds);
  const float k = 2.f * g[b] / per_sample;
  const int64_t base = (int64_t)b * per_sample;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < per_sample; i += gridDim.x * blockDim.x * 4) {
    F4 p = load4(pred + base + i), xv = load4(x + base + i), ev = load4(eps + base + i), o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = k * (p.v[j] - (da * xv.v[j] + ds * ev.v[j]));
    store4(dpred + base + i, o);
  }
}

// One warp per token row: dot, |z~|, |z| -> stats[row] = {dot, n1, n2}; align[b] += -(dot/(n1 n2)) / T
template <typename TP, typename TZ>
__global__ void __launch_bounds__(128) cos_fwd_kernel(const TP* __restrict__ zt, const TZ* __restrict__ z,
                                                       float* __restrict__ stats, float* __restrict__ align,
                                                       int64_t rows, int T, int Z) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TP* a = zt + row * Z;
  const TZ* b = z + row * Z;
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (int c = lane * 4; c < Z; c += 128) {
    F4 u = load4(a + c), v = load4(b + c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dot += u.v[j] * v.v[j];
      na += u.v[j] * u.v[j];
      nb += v.v[j] * v.v[j];
    }
  }
  dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
  if (lane == 0) {
    float n1 = fmaxf(sqrtf(na), 1e-12f), n2 = fmaxf(sqrtf(nb), 1e-12f);
    stats[row * 3 + 0] = dot;
    stats[row * 3 + 1] = n1;
    stats[row * 3 + 2] = n2;
    atomicAdd(align + row / T, -(dot / (n1 * n2)) / T);
  }
}

// dz~ = g[b] * (-1/T) * ( z/(n1 n2) - dot z~/(n1^3 n2) )     (n1 clamped: gradient of the clamp branch is z/(eps n2))
template <typename TP, typename TZ>
__global__ void __launch_bounds__(128) cos_bwd_kernel(const TP* __restrict__ zt, const TZ* __restrict__ z,
                                                       const float* __restrict__ stats, const float* __restrict__ g,
                                                       TP* __restrict__ dzt, int64_t rows, int T, int Z) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float dot = stats[row * 3], n1 = stats[row * 3 + 1], n2 = stats[row * 3 + 2];
  const float k = -g[row / T] / T;
  const float c_z = k / (n1 * n2);
  const float c_a = n1 > 1e-12f ? -k * dot / (n1 * n1 * n1 * n2) : 0.f;
  const TP* a = zt + row * Z;
  const TZ* b = z + row * Z;
  TP* o = dzt + row * Z;
  for (int c = lane * 4; c < Z; c += 128) {
    F4 u = load4(a + c), v = load4(b + c), r;
#pragma unroll
    for (int j = 0; j < 4; ++j) r.v[j] = c_z * v.v[j] + c_a * u.v[j];
    store4(o + c, r);
  }
}

}  // namespace reed

using namespace reed;

extern "C" int reed_siloss_interp(const void* x, const void* eps, const void* t, void* xt, int batch, int per_sample,
                                  int path_type, void* stream) {
  REED_REQUIRE(per_sample % 4 == 0, "siloss: per-sample size must be a multiple of 4");
  REED_REQUIRE(path_type == 0 || path_type == 1, "siloss: path_type 0 (linear) or 1 (cosine)");
  if (batch == 0) return 0;
  dim3 grid(ceil_div(per_sample, 1024) > 8 ? 8 : ceil_div(per_sample, 1024), batch);
  interp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)eps, (const float*)t, (float*)xt,
                                                       per_sample, path_type);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_siloss_mse_fwd(const void* pred, const void* x, const void* eps, const void* t, void* denoise,
                                   int batch, int per_sample, int path_type, void* stream) {
  REED_REQUIRE(per_sample % 4 == 0, "siloss: per-sample size must be a multiple of 4");
  if (batch == 0) return 0;
  mse_fwd_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>((const float*)pred, (const float*)x, (const float*)eps,
                                                         (const float*)t, (float*)denoise, per_sample, path_type);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_siloss_mse_bwd(const void* pred, const void* x, const void* eps, const void* t, const void* g,
                                   void* dpred, int batch, int per_sample, int path_type, void* stream) {
  REED_REQUIRE(per_sample % 4 == 0, "siloss: per-sample size must be a multiple of 4");
  if (batch == 0) return 0;
  dim3 grid(ceil_div(per_sample, 1024) > 8 ? 8 : ceil_div(per_sample, 1024), batch);
  mse_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)pred, (const float*)x, (const float*)eps,
                                                        (const float*)t, (const float*)g, (float*)dpred, per_sample,
                                                        path_type);
  REED_LAUNCH_CHECK();
  return 0;
}

// align must be zero-initialised by the caller; stats is [rows, 3] fp32
extern "C" int reed_siloss_cos_fwd(const void* zt, int zt_dtype, const void* z, int z_dtype, void* stats, void* align,
                                   int batch, int T, int Z, void* stream) {
  REED_REQUIRE(Z % 4 == 0, "siloss: feature width must be a multiple of 4");
  const int64_t rows = (int64_t)batch * T;
  if (rows == 0) return 0;
  dim3 grid(ceil_div(rows, 4));
  cudaStream_t st = (cudaStream_t)stream;
#define CF(TP, TZ) cos_fwd_kernel<TP, TZ><<<grid, 128, 0, st>>>((const TP*)zt, (const TZ*)z, (float*)stats, (float*)align, rows, T, Z)
  if (zt_dtype == kF32 && z_dtype == kF32) CF(float, float);
  else if (zt_dtype == kF32) CF(float, bf16);
  else if (z_dtype == kF32) CF(bf16, float);
  else CF(bf16, bf16);
#undef CF
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_siloss_cos_bwd(const void* zt, int zt_dtype, const void* z, int z_dtype, const void* stats,
                                   const void* g, void* dzt, int batch, int T, int Z, void* stream) {
  REED_REQUIRE(Z % 4 == 0, "siloss: feature width must be a multiple of 4");
  const int64_t rows = (int64_t)batch * T;
  if (rows == 0) return 0;
  dim3 grid(ceil_div(rows, 4));
  cudaStream_t st = (cudaStream_t)stream;
#define CB(TP, TZ) cos_bwd_kernel<TP, TZ><<<grid, 128, 0, st>>>((const TP*)zt, (const TZ*)z, (const float*)stats, (const float*)g, (TP*)dzt, rows, T, Z)
  if (zt_dtype == kF32 && z_dtype == kF32) CB(float, float);
  else if (zt_dtype == kF32) CB(float, bf16);
  else if (z_dtype == kF32) CB(bf16, float);
  else CB(bf16, bf16);
#undef CB
  REED_LAUNCH_CHECK();
  return 0;
}
