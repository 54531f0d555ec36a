// Fused sampler-step kernels (north_star kernel 5): fp64 state, model evaluated in fp32/bf16.
// One pass per step does the cast-in of the model output, score-from-velocity, drift, classifier-free-guidance
// combine, the Euler / Heun / Euler-Maruyama update and the cast-out (optionally batch-duplicated for CFG) of
// the next model input.
//
// Reference: /root/reference/image/samplers.py:15-39 (score), 42-43 (diffusion w = 2t), 61-104 (ODE Euler/Heun),
// 124-187 (SDE Euler-Maruyama; the last step is the deterministic mean update).
#include "common.cuh"

namespace reed {

struct StepArgs {
  const double* x_cur;    // [n]
  const void* v;          // model output, [n] or [2n] (cond half first) in the model dtype
  const double* eps;      // SDE noise [n] or null
  const double* d_prev;   // Heun stage 2: guided slope of stage 1, else null
  double* d_out;          // optional: guided slope of this evaluation
  double* x_next;         // [n]
  void* x_model;          // optional cast of x_next in the model dtype, [n] or [2n]
  int64_t n;              // elements per batch half
  int guided;             // v holds 2n elements
  int dup_out;            // write x_model twice (next evaluation is guided)
  int sde;                // 0: slope = v ; 1: slope = v - 0.5 w s
  int path;               // 0 linear, 1 cosine
  double cfg, t_cur, dt;
};

template <typename TM>
__global__ void __launch_bounds__(256) sampler_step_kernel(StepArgs a) {
  // per-step scalars (fp64, same expressions as the reference)
  double ratio = 0.0, var = 1.0;
  if (a.sde) {
    if (a.path == 0) {
      ratio = (1.0 - a.t_cur) / -1.0;
      var = a.t_cur * a.t_cur - ratio * 1.0 * a.t_cur;
    } else {
      const double hp = 1.5707963267948966;
      const double al = cos(a.t_cur * hp), si = sin(a.t_cur * hp);
      ratio = al / (-hp * si);
      var = si * si - ratio * (hp * al) * si;
    }
  }
  const double w = 2.0 * a.t_cur;
  const double noise_scale = sqrt(w) * sqrt(fabs(a.dt));
  const TM* v = reinterpret_cast<const TM*>(a.v);
  TM* xm = reinterpret_cast<TM*>(a.x_model);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = a.x_cur[i];
    double d = (double)to_f(v[i]);
    if (a.sde) d = d - 0.5 * w * ((ratio * d - x) / var);
    if (a.guided) {
      double du = (double)to_f(v[a.n + i]);
      if (a.sde) du = du - 0.5 * w * ((ratio * du - x) / var);
      d = du + a.cfg * (d - du);
    }
    if (a.d_out) a.d_out[i] = d;
    double xn;
    if (a.d_prev) xn = x + a.dt * (0.5 * a.d_prev[i] + 0.5 * d);
    else xn = x + d * a.dt;
    if (a.eps) xn = xn + noise_scale * a.eps[i];
    a.x_next[i] = xn;
    if (xm) {
      TM c = from_f<TM>((float)xn);
      xm[i] = c;
      if (a.dup_out) xm[a.n + i] = c;
    }
  }
}

// x_model = cast(x) (optionally duplicated): the first evaluation of a sampling run
template <typename TM>
__global__ void __launch_bounds__(256) sampler_cast_kernel(const double* __restrict__ x, TM* __restrict__ xm, int64_t n,
                                                            int dup) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    TM c = from_f<TM>((float)x[i]);
    xm[i] = c;
    if (dup) xm[n + i] = c;
  }
}

}  // namespace reed

using namespace reed;

static inline int step_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > kNumSMs * 8 ? kNumSMs * 8 : b));
}

extern "C" int reed_sampler_step(const void* x_cur, const void* v, int model_dtype, const void* eps, const void* d_prev,
                                 void* d_out, void* x_next, void* x_model, int64_t n, int guided, int dup_out, int sde,
                                 int path_type, double cfg, double t_cur, double dt, void* stream) {
  REED_REQUIRE(path_type == 0 || path_type == 1, "sampler: path_type 0 (linear) or 1 (cosine)");
  if (n == 0) return 0;
  StepArgs a{(const double*)x_cur, v, (const double*)eps, (const double*)d_prev, (double*)d_out, (double*)x_next, x_model,
             n, guided, dup_out, sde, path_type, cfg, t_cur, dt};
  if (model_dtype == kBF16) sampler_step_kernel<bf16><<<step_grid(n), 256, 0, (cudaStream_t)stream>>>(a);
  else sampler_step_kernel<float><<<step_grid(n), 256, 0, (cudaStream_t)stream>>>(a);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_sampler_cast(const void* x, void* x_model, int model_dtype, int64_t n, int dup, void* stream) {
  if (n == 0) return 0;
  if (model_dtype == kBF16) sampler_cast_kernel<bf16><<<step_grid(n), 256, 0, (cudaStream_t)stream>>>((const double*)x, (bf16*)x_model, n, dup);
  else sampler_cast_kernel<float><<<step_grid(n), 256, 0, (cudaStream_t)stream>>>((const double*)x, (float*)x_model, n, dup);
  REED_LAUNCH_CHECK();
  return 0;
}
