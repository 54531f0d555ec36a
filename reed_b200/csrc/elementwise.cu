// HBM-bound fused kernels of the SiT block: adaLN-Zero LayerNorm+modulate (fwd/bwd), gate backward,
// casts/activations, bias-gradient column sums, token mean.  One warp owns one token row; rows are cached
// in registers (<= 16 float4 per lane, D <= 2048) so every tensor is read exactly once.
//
// Reference semantics: /root/reference/image/models/sit.py:26-27 (modulate), 113/119/146 (LayerNorm, no
// affine, eps 1e-6), 130-137 (block), 153-158 (final layer).
#include "common.cuh"

namespace reed {

constexpr int kRowWarps = 4;   // warps per CTA in the row kernels

// ---------------------------------------------------------------------------------------------
// out[m,:] = LN(x[m,:]) * (1 + scale[g,:]) + shift[g,:],  g = m / rows_per_group
// ---------------------------------------------------------------------------------------------
template <typename TA, int MAXV>
__global__ void __launch_bounds__(kRowWarps * 32) ln_modulate_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale, int64_t ld_mod,
    int rows_per_group, TA* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D,
    float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + (int64_t)row * D;
  F4 c[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int col = (i * 32 + lane) * 4;
    if (col < D) {
      c[i] = load4(xr + col);
      s += (c[i].v[0] + c[i].v[1]) + (c[i].v[2] + c[i].v[3]);
    }
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int col = (i * 32 + lane) * 4;
    if (col < D) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float d = c[i].v[j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const int g = row / rows_per_group;
  const float* sh = shift + (int64_t)g * ld_mod;
  const float* sc = scale + (int64_t)g * ld_mod;
  TA* o = out + (int64_t)row * D;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int col = (i * 32 + lane) * 4;
    if (col < D) {
      F4 a = load4(sh + col), b = load4(sc + col), r;
#pragma unroll
      for (int j = 0; j < 4; ++j) r.v[j] = (c[i].v[j] - mean) * rstd * (1.f + b.v[j]) + a.v[j];
      store4(o + col, r);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of the above, fused with the residual-gradient add:
//   dx[m,:] = dres[m,:] + rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat)),  dxhat = dout*(1+scale)
//   dshift[g,:] += sum_m dout ; dscale[g,:] += sum_m dout*xhat        (fp32 atomics, one per CTA per column)
// Each CTA owns `rows_per_cta` consecutive rows of ONE group.
// ---------------------------------------------------------------------------------------------
template <typename TA, int MAXV>
__global__ void __launch_bounds__(kRowWarps * 32) ln_modulate_bwd_kernel(
    const TA* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ mean_in,
    const float* __restrict__ rstd_in, const float* __restrict__ scale, int64_t ld_mod, int rows_per_group,
    const float* __restrict__ dres, float* __restrict__ dx, float* __restrict__ dshift, float* __restrict__ dscale,
    int M, int D, int rows_per_cta) {
  extern __shared__ float red[];   // [2][kRowWarps][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks_per_group = (rows_per_group + rows_per_cta - 1) / rows_per_cta;
  const int g = blockIdx.x / chunks_per_group;
  const int r0 = g * rows_per_group + (blockIdx.x % chunks_per_group) * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, (g + 1) * rows_per_group);
  const float* sc = scale + (int64_t)g * ld_mod;

  F4 one_plus[MAXV], acc_sh[MAXV], acc_sc[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int col = (i * 32 + lane) * 4;
    acc_sh[i] = F4{{0, 0, 0, 0}};
    acc_sc[i] = F4{{0, 0, 0, 0}};
    one_plus[i] = F4{{1, 1, 1, 1}};
    if (col < D) {
      F4 s = load4(sc + col);
#pragma unroll
      for (int j = 0; j < 4; ++j) one_plus[i].v[j] = 1.f + s.v[j];
    }
  }
  for (int row = r0 + warp; row < r1; row += kRowWarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    const TA* dor = dout + (int64_t)row * D;
    const float* xr = x + (int64_t)row * D;
    F4 xh[MAXV], dh[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int col = (i * 32 + lane) * 4;
      if (col < D) {
        F4 d = load4(dor + col);
        xh[i] = load4(xr + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float h = (xh[i].v[j] - mean) * rstd;
          xh[i].v[j] = h;
          acc_sh[i].v[j] += d.v[j];
          acc_sc[i].v[j] += d.v[j] * h;
          float e = d.v[j] * one_plus[i].v[j];
          dh[i].v[j] = e;
          s1 += e;
          s2 += e * h;
        }
      }
    }
    const float c1 = warp_sum(s1) / D, c2 = warp_sum(s2) / D;
    float* dxr = dx + (int64_t)row * D;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int col = (i * 32 + lane) * 4;
      if (col < D) {
        F4 r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = rstd * (dh[i].v[j] - c1 - xh[i].v[j] * c2);
        if (dres != nullptr) {
          F4 p = load4(dres + (int64_t)row * D + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) r.v[j] += p.v[j];
        }
        store4(dxr + col, r);
      }
    }
  }
  // CTA-level reduce of the column partials, then one atomic per column
  float* red_sh = red + (int64_t)warp * D;
  float* red_sc = red + (int64_t)(kRowWarps + warp) * D;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int col = (i * 32 + lane) * 4;
    if (col < D) {
      store4(red_sh + col, acc_sh[i]);
      store4(red_sc + col, acc_sc[i]);
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < D; col += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) {
      a += red[(int64_t)w * D + col];
      b += red[(int64_t)(kRowWarps + w) * D + col];
    }
    atomicAdd(dshift + (int64_t)g * ld_mod + col, a);
    atomicAdd(dscale + (int64_t)g * ld_mod + col, b);
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of x_new = x + gate * y:   dy = gate * dxn (act dtype),  dgate[g,:] += sum_m dxn*y,
// and (optional) dbias[:] += sum_m dy  (y = GEMM + bias, so d bias = column sum of dy).
// ---------------------------------------------------------------------------------------------
template <typename TA, int MAXV>
__global__ void __launch_bounds__(kRowWarps * 32) gate_bwd_kernel(
    const float* __restrict__ dxn, const TA* __restrict__ y, const float* __restrict__ gate, int64_t ld_mod,
    int rows_per_group, TA* __restrict__ dy, float* __restrict__ dgate, float* __restrict__ dbias, int M, int D,
    int rows_per_cta) {
  extern __shared__ float red[];   // [2][kRowWarps][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks_per_group = (rows_per_group + rows_per_cta - 1) / rows_per_cta;
  const int g = blockIdx.x / chunks_per_group;
  const int r0 = g * rows_per_group + (blockIdx.x % chunks_per_group) * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, (g + 1) * rows_per_group);
  const float* gt = gate + (int64_t)g * ld_mod;
  F4 gv[MAXV], acc_g[MAXV], acc_b[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int col = (i * 32 + lane) * 4;
    acc_g[i] = F4{{0, 0, 0, 0}};
    acc_b[i] = F4{{0, 0, 0, 0}};
    gv[i] = F4{{0, 0, 0, 0}};
    if (col < D) gv[i] = load4(gt + col);
  }
  for (int row = r0 + warp; row < r1; row += kRowWarps) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int col = (i * 32 + lane) * 4;
      if (col < D) {
        F4 d = load4(dxn + (int64_t)row * D + col);
        F4 yy = load4(y + (int64_t)row * D + col);
        F4 o;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc_g[i].v[j] += d.v[j] * yy.v[j];
          o.v[j] = to_f(from_f<TA>(d.v[j] * gv[i].v[j]));
          acc_b[i].v[j] += o.v[j];
        }
        store4(dy + (int64_t)row * D + col, o);
      }
    }
  }
  float* red_g = red + (int64_t)warp * D;
  float* red_b = red + (int64_t)(kRowWarps + warp) * D;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int col = (i * 32 + lane) * 4;
    if (col < D) {
      store4(red_g + col, acc_g[i]);
      store4(red_b + col, acc_b[i]);
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < D; col += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) {
      a += red[(int64_t)w * D + col];
      b += red[(int64_t)(kRowWarps + w) * D + col];
    }
    atomicAdd(dgate + (int64_t)g * ld_mod + col, a);
    if (dbias != nullptr) atomicAdd(dbias + col, b);
  }
}

// ---------------------------------------------------------------------------------------------
// column sums: out[n] (+)= sum_m src[m, n]      (bias gradients)
// ---------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256) colsum_kernel(const TA* __restrict__ src, int64_t ld, float* __restrict__ out,
                                                      int M, int N, int rows_per_cta) {
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, M);
  F4 acc{{0, 0, 0, 0}};
  if (col < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      F4 v = load4(src + (int64_t)r * ld + col);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc.v[j] += v.v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[warp][lane * 4 + j] = acc.v[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    int c = blockIdx.x * 128 + threadIdx.x;
    if (c < N) atomicAdd(out + c, s);
  }
}

// ---------------------------------------------------------------------------------------------
// flat elementwise: casts, SiLU fwd/bwd, act-derivative multiply
// ---------------------------------------------------------------------------------------------
enum UnaryOp { kCast = 0, kSilu = 1 };

template <typename TI, typename TO, int OP>
__global__ void __launch_bounds__(256) unary_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 v = load4(in + i * 4);
    if (OP == kSilu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v.v[j] = silu(v.v[j]);
    }
    store4(out + i * 4, v);
  }
}

// dx = dy * f'(h), f in {silu, gelu_tanh}; TH = dtype of h, TD = dtype of dy/dx
template <typename TH, typename TD, int GELU>
__global__ void __launch_bounds__(256) act_bwd_kernel(const TD* __restrict__ dy, const TH* __restrict__ h,
                                                       TD* __restrict__ dx, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 d = load4(dy + i * 4), hv = load4(h + i * 4), o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = d.v[j] * (GELU ? gelu_tanh_grad(hv.v[j]) : silu_grad(hv.v[j]));
    store4(dx + i * 4, o);
  }
}

// out[g, :] = mean over the group's rows of x (token mean for the 't' projector); bwd broadcasts dy/T
template <typename TO>
__global__ void __launch_bounds__(256) group_mean_kernel(const float* __restrict__ x, TO* __restrict__ out,
                                                          int rows_per_group, int D) {
  const int g = blockIdx.y;
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (col >= D) return;
  F4 acc{{0, 0, 0, 0}};
  const float* base = x + (int64_t)g * rows_per_group * D + col;
  for (int r = 0; r < rows_per_group; ++r) {
    F4 v = load4(base + (int64_t)r * D);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.v[j] += v.v[j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) acc.v[j] /= rows_per_group;
  store4(out + (int64_t)g * D + col, acc);
}

// dx[m,:] (+)= dy[g,:] / T
__global__ void __launch_bounds__(256) group_mean_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                              int rows_per_group, int D, int64_t n4, int accumulate) {
  const int d4 = D / 4;
  const float inv = 1.f / rows_per_group;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / d4;
    int col = (int)(i - row * d4) * 4;
    F4 v = load4(dy + (row / rows_per_group) * D + col);
    F4 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = v.v[j] * inv;
    if (accumulate) {
      F4 p = load4(dx + i * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) o.v[j] += p.v[j];
    }
    store4(dx + i * 4, o);
  }
}

// out = a + b (fp32), used to merge the projector-tap gradient into the residual-stream gradient
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   float* __restrict__ out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 x = load4(a + i * 4), y = load4(b + i * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) x.v[j] += y.v[j];
    store4(out + i * 4, x);
  }
}

static inline int flat_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

template <typename TA>
static int ln_fwd_dispatch(const float* x, const float* shift, const float* scale, int64_t ld_mod, int rpg, void* out,
                           float* mean, float* rstd, int M, int D, float eps, cudaStream_t st) {
  dim3 grid(ceil_div(M, kRowWarps)), block(kRowWarps * 32);
  int nv = ceil_div(D, 128);
#define LN_FWD(V) ln_modulate_fwd_kernel<TA, V><<<grid, block, 0, st>>>(x, shift, scale, ld_mod, rpg, (TA*)out, mean, rstd, M, D, eps)
  if (nv <= 4) LN_FWD(4); else if (nv <= 8) LN_FWD(8); else if (nv <= 12) LN_FWD(12); else LN_FWD(16);
#undef LN_FWD
  REED_LAUNCH_CHECK();
  return 0;
}

template <typename TA>
static int ln_bwd_dispatch(const void* dout, const float* x, const float* mean, const float* rstd, const float* scale,
                           int64_t ld_mod, int rpg, const float* dres, float* dx, float* dshift, float* dscale, int M,
                           int D, cudaStream_t st) {
  int rows_per_cta = rpg < 32 ? rpg : 32;
  int chunks = ceil_div(rpg, rows_per_cta);
  dim3 grid((M / rpg) * chunks), block(kRowWarps * 32);
  size_t smem = sizeof(float) * 2 * kRowWarps * D;
  int nv = ceil_div(D, 128);
#define LN_BWD(V) ln_modulate_bwd_kernel<TA, V><<<grid, block, smem, st>>>((const TA*)dout, x, mean, rstd, scale, ld_mod, rpg, dres, dx, dshift, dscale, M, D, rows_per_cta)
  if (nv <= 4) LN_BWD(4); else if (nv <= 8) LN_BWD(8); else if (nv <= 12) LN_BWD(12); else LN_BWD(16);
#undef LN_BWD
  REED_LAUNCH_CHECK();
  return 0;
}

template <typename TA>
static int gate_bwd_dispatch(const float* dxn, const void* y, const float* gate, int64_t ld_mod, int rpg, void* dy,
                             float* dgate, float* dbias, int M, int D, cudaStream_t st) {
  int rows_per_cta = rpg < 32 ? rpg : 32;
  int chunks = ceil_div(rpg, rows_per_cta);
  dim3 grid((M / rpg) * chunks), block(kRowWarps * 32);
  size_t smem = sizeof(float) * 2 * kRowWarps * D;
  int nv = ceil_div(D, 128);
#define GB(V) gate_bwd_kernel<TA, V><<<grid, block, smem, st>>>(dxn, (const TA*)y, gate, ld_mod, rpg, (TA*)dy, dgate, dbias, M, D, rows_per_cta)
  if (nv <= 4) GB(4); else if (nv <= 8) GB(8); else if (nv <= 12) GB(12); else GB(16);
#undef GB
  REED_LAUNCH_CHECK();
  return 0;
}

}  // namespace reed

using namespace reed;

#define ROW_ARGS_OK(M, D, rpg)                                                                   \
  REED_REQUIRE((D) % 4 == 0 && (D) <= 2048, "row kernels need D %% 4 == 0 and D <= 2048, got %d", (int)(D)); \
  REED_REQUIRE((rpg) > 0 && (M) % (rpg) == 0, "M=%d not a multiple of rows_per_group=%d", (int)(M), (int)(rpg))

extern "C" int reed_ln_modulate_fwd(const void* x, const void* shift, const void* scale, int64_t ld_mod,
                                    int rows_per_group, void* out, int act_dtype, void* mean, void* rstd, int M, int D,
                                    float eps, void* stream) {
  ROW_ARGS_OK(M, D, rows_per_group);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16)
    return ln_fwd_dispatch<bf16>((const float*)x, (const float*)shift, (const float*)scale, ld_mod, rows_per_group, out,
                                 (float*)mean, (float*)rstd, M, D, eps, st);
  return ln_fwd_dispatch<float>((const float*)x, (const float*)shift, (const float*)scale, ld_mod, rows_per_group, out,
                                (float*)mean, (float*)rstd, M, D, eps, st);
}

extern "C" int reed_ln_modulate_bwd(const void* dout, int act_dtype, const void* x, const void* mean, const void* rstd,
                                    const void* scale, int64_t ld_mod, int rows_per_group, const void* dres, void* dx,
                                    void* dshift, void* dscale, int M, int D, void* stream) {
  ROW_ARGS_OK(M, D, rows_per_group);
  REED_REQUIRE(D <= 1536, "ln_modulate_bwd: D <= 1536 (48 KB reduction buffer), got %d", D);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16)
    return ln_bwd_dispatch<bf16>(dout, (const float*)x, (const float*)mean, (const float*)rstd, (const float*)scale,
                                 ld_mod, rows_per_group, (const float*)dres, (float*)dx, (float*)dshift, (float*)dscale,
                                 M, D, st);
  return ln_bwd_dispatch<float>(dout, (const float*)x, (const float*)mean, (const float*)rstd, (const float*)scale,
                                ld_mod, rows_per_group, (const float*)dres, (float*)dx, (float*)dshift, (float*)dscale, M,
                                D, st);
}

extern "C" int reed_gate_bwd(const void* dxn, const void* y, int act_dtype, const void* gate, int64_t ld_mod,
                             int rows_per_group, void* dy, void* dgate, void* dbias, int M, int D, void* stream) {
  ROW_ARGS_OK(M, D, rows_per_group);
  REED_REQUIRE(D <= 1536, "gate_bwd: D <= 1536 (48 KB reduction buffer), got %d", D);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16)
    return gate_bwd_dispatch<bf16>((const float*)dxn, y, (const float*)gate, ld_mod, rows_per_group, dy, (float*)dgate,
                                   (float*)dbias, M, D, st);
  return gate_bwd_dispatch<float>((const float*)dxn, y, (const float*)gate, ld_mod, rows_per_group, dy, (float*)dgate,
                                  (float*)dbias, M, D, st);
}

extern "C" int reed_colsum(const void* src, int act_dtype, int64_t ld, void* out, int M, int N, void* stream) {
  REED_REQUIRE(N % 4 == 0 && ld % 4 == 0, "colsum needs N, ld %% 4 == 0");
  if (M == 0 || N == 0) return 0;
  int rows_per_cta = 256;
  dim3 grid(ceil_div(N, 128), ceil_div(M, rows_per_cta));
  if (act_dtype == kBF16)
    colsum_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, ld, (float*)out, M, N, rows_per_cta);
  else
    colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, ld, (float*)out, M, N, rows_per_cta);
  REED_LAUNCH_CHECK();
  return 0;
}

// op: 0 cast, 1 silu.  in_dtype/out_dtype: 0 fp32, 1 bf16 (fp32->fp32, fp32->bf16, bf16->fp32 supported)
extern "C" int reed_unary(const void* in, int in_dtype, void* out, int out_dtype, int op, int64_t n, void* stream) {
  REED_REQUIRE(n % 4 == 0, "unary needs n %% 4 == 0, got %lld", (long long)n);
  if (n == 0) return 0;
  int64_t n4 = n / 4;
  int grid = flat_grid(n4);
  cudaStream_t st = (cudaStream_t)stream;
#define UN(TI, TO, OP) unary_kernel<TI, TO, OP><<<grid, 256, 0, st>>>((const TI*)in, (TO*)out, n4)
  if (in_dtype == kF32 && out_dtype == kF32) { if (op == kSilu) UN(float, float, kSilu); else UN(float, float, kCast); }
  else if (in_dtype == kF32 && out_dtype == kBF16) { if (op == kSilu) UN(float, bf16, kSilu); else UN(float, bf16, kCast); }
  else if (in_dtype == kBF16 && out_dtype == kF32) { if (op == kSilu) UN(bf16, float, kSilu); else UN(bf16, float, kCast); }
  else return fail("reed_unary: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
#undef UN
  REED_LAUNCH_CHECK();
  return 0;
}

// dx = dy * act'(h); act: 1 gelu_tanh, 2 silu.  h_dtype: dtype of h; d_dtype: dtype of dy/dx
extern "C" int reed_act_bwd(const void* dy, int d_dtype, const void* h, int h_dtype, void* dx, int act, int64_t n,
                            void* stream) {
  REED_REQUIRE(n % 4 == 0, "act_bwd needs n %% 4 == 0");
  if (n == 0) return 0;
  int64_t n4 = n / 4;
  int grid = flat_grid(n4);
  cudaStream_t st = (cudaStream_t)stream;
  const bool gelu = act == 1;
#define AB(TH, TD) do { if (gelu) act_bwd_kernel<TH, TD, 1><<<grid, 256, 0, st>>>((const TD*)dy, (const TH*)h, (TD*)dx, n4); \
                        else act_bwd_kernel<TH, TD, 0><<<grid, 256, 0, st>>>((const TD*)dy, (const TH*)h, (TD*)dx, n4); } while (0)
  if (h_dtype == kF32 && d_dtype == kF32) AB(float, float);
  else if (h_dtype == kBF16 && d_dtype == kBF16) AB(bf16, bf16);
  else if (h_dtype == kF32 && d_dtype == kBF16) AB(float, bf16);
  else AB(bf16, float);
#undef AB
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_group_mean_fwd(const void* x, void* out, int out_dtype, int groups, int rows_per_group, int D,
                                   void* stream) {
  REED_REQUIRE(D % 4 == 0, "group_mean needs D %% 4 == 0");
  if (groups == 0) return 0;
  dim3 grid(ceil_div(D / 4, 256), groups);
  if (out_dtype == kBF16)
    group_mean_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (bf16*)out, rows_per_group, D);
  else
    group_mean_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)out, rows_per_group, D);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_group_mean_bwd(const void* dy, void* dx, int groups, int rows_per_group, int D, int accumulate,
                                   void* stream) {
  REED_REQUIRE(D % 4 == 0, "group_mean needs D %% 4 == 0");
  int64_t n4 = (int64_t)groups * rows_per_group * D / 4;
  if (n4 == 0) return 0;
  group_mean_bwd_kernel<<<flat_grid(n4), 256, 0, (cudaStream_t)stream>>>((const float*)dy, (float*)dx, rows_per_group, D,
                                                                        n4, accumulate);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_add_f32(const void* a, const void* b, void* out, int64_t n, void* stream) {
  REED_REQUIRE(n % 4 == 0, "add needs n %% 4 == 0");
  if (n == 0) return 0;
  add_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float*)a, (const float*)b, (float*)out, n / 4);
  REED_LAUNCH_CHECK();
  return 0;
}
