// fp32-math multi-head attention (non-causal, no mask, no dropout) over the packed QKV buffer produced by the
// qkv GEMM: qkv[b, t, {q,k,v}, h, d]  ->  o[b, t, h, d].  This is the attention of the fp32 precision mode
// (parity bar 1e-5) and the fallback for head sizes the tensor-core kernel does not take.  One warp per
// query row (fwd, dQ) or per key row (dK, dV); no atomics, deterministic.
//
// Reference semantics: timm Attention (imported at /root/reference/image/models/sit.py:13, used 114-118,134):
// softmax(q k^T * d^-0.5) v with q,k,v = qkv.reshape(B,N,3,H,d).permute(2,0,3,1,4).
#include "common.cuh"

namespace reed {

constexpr int kAttWarps = 4;
constexpr int kMaxHD = 128;

template <typename TA>
__global__ void __launch_bounds__(kAttWarps * 32) attn_simt_fwd_kernel(const TA* __restrict__ qkv, TA* __restrict__ o,
                                                                         float* __restrict__ lse, int B, int T, int H,
                                                                         int hd, float scale) {
  extern __shared__ float sm[];   // per warp: scores[T] + q[hd]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kAttWarps + warp;   // over B*H*T, t fastest
  if (row >= (int64_t)B * H * T) return;
  const int t = (int)(row % T);
  const int h = (int)((row / T) % H);
  const int b = (int)(row / ((int64_t)T * H));
  const int64_t tok_stride = 3LL * H * hd;
  float* sc = sm + (int64_t)warp * (T + kMaxHD);
  float* qs = sc + T;
  const TA* qp = qkv + ((int64_t)b * T + t) * tok_stride + h * hd;
  for (int d = lane; d < hd; d += 32) qs[d] = to_f(qp[d]) * scale;
  __syncwarp();
  const TA* kbase = qkv + (int64_t)b * T * tok_stride + (int64_t)H * hd + h * hd;
  const TA* vbase = kbase + (int64_t)H * hd;
  float mx = -INFINITY;
  for (int j = lane; j < T; j += 32) {
    const TA* kp = kbase + (int64_t)j * tok_stride;
    float s = 0.f;
    for (int d = 0; d < hd; d += 4) {
      F4 kv = load4(kp + d);
      s += qs[d] * kv.v[0] + qs[d + 1] * kv.v[1] + qs[d + 2] * kv.v[2] + qs[d + 3] * kv.v[3];
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) {
    float p = expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  if (lane == 0) lse[row] = mx + logf(sum);
  float acc[kMaxHD / 32];
#pragma unroll
  for (int i = 0; i < kMaxHD / 32; ++i) acc[i] = 0.f;
  for (int j = 0; j < T; ++j) {
    const float p = sc[j];
    const TA* vp = vbase + (int64_t)j * tok_stride;
#pragma unroll
    for (int i = 0; i < kMaxHD / 32; ++i) {
      int d = lane + i * 32;
      if (d < hd) acc[i] += p * to_f(vp[d]);
    }
  }
  TA* op = o + ((int64_t)b * T + t) * ((int64_t)H * hd) + h * hd;
#pragma unroll
  for (int i = 0; i < kMaxHD / 32; ++i) {
    int d = lane + i * 32;
    if (d < hd) op[d] = from_f<TA>(acc[i] * inv);
  }
}

// dQ and the row statistic delta_i = dO_i . O_i
template <typename TA>
__global__ void __launch_bounds__(kAttWarps * 32) attn_simt_bwd_q_kernel(
    const TA* __restrict__ qkv, const TA* __restrict__ o, const TA* __restrict__ d_o, const float* __restrict__ lse,
    TA* __restrict__ dqkv, float* __restrict__ delta, int B, int T, int H, int hd, float scale) {
  extern __shared__ float sm[];   // per warp: ds[T] + q[hd] + do[hd]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kAttWarps + warp;
  if (row >= (int64_t)B * H * T) return;
  const int t = (int)(row % T);
  const int h = (int)((row / T) % H);
  const int b = (int)(row / ((int64_t)T * H));
  const int64_t tok_stride = 3LL * H * hd;
  float* ds = sm + (int64_t)warp * (T + 2 * kMaxHD);
  float* qs = ds + T;
  float* dos = qs + kMaxHD;
  const TA* qp = qkv + ((int64_t)b * T + t) * tok_stride + h * hd;
  const int64_t orow = ((int64_t)b * T + t) * ((int64_t)H * hd) + h * hd;
  float dl = 0.f;
  for (int d = lane; d < hd; d += 32) {
    qs[d] = to_f(qp[d]) * scale;
    float g = to_f(d_o[orow + d]);
    dos[d] = g;
    dl += g * to_f(o[orow + d]);
  }
  dl = warp_sum(dl);
  if (lane == 0) delta[row] = dl;
  __syncwarp();
  const float l = lse[row];
  const TA* kbase = qkv + (int64_t)b * T * tok_stride + (int64_t)H * hd + h * hd;
  const TA* vbase = kbase + (int64_t)H * hd;
  for (int j = lane; j < T; j += 32) {
    const TA* kp = kbase + (int64_t)j * tok_stride;
    const TA* vp = vbase + (int64_t)j * tok_stride;
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < hd; d += 4) {
      F4 kv = load4(kp + d), vv = load4(vp + d);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s += qs[d + c] * kv.v[c];
        dp += dos[d + c] * vv.v[c];
      }
    }
    ds[j] = expf(s - l) * (dp - dl);
  }
  __syncwarp();
  float acc[kMaxHD / 32];
#pragma unroll
  for (int i = 0; i < kMaxHD / 32; ++i) acc[i] = 0.f;
  for (int j = 0; j < T; ++j) {
    const float w = ds[j];
    const TA* kp = kbase + (int64_t)j * tok_stride;
#pragma unroll
    for (int i = 0; i < kMaxHD / 32; ++i) {
      int d = lane + i * 32;
      if (d < hd) acc[i] += w * to_f(kp[d]);
    }
  }
  TA* dq = dqkv + ((int64_t)b * T + t) * tok_stride + h * hd;
#pragma unroll
  for (int i = 0; i < kMaxHD / 32; ++i) {
    int d = lane + i * 32;
    if (d < hd) dq[d] = from_f<TA>(acc[i] * scale);
  }
}

// dK and dV: one warp per key row j
template <typename TA>
__global__ void __launch_bounds__(kAttWarps * 32) attn_simt_bwd_kv_kernel(
    const TA* __restrict__ qkv, const TA* __restrict__ d_o, const float* __restrict__ lse,
    const float* __restrict__ delta, TA* __restrict__ dqkv, int B, int T, int H, int hd, float scale) {
  extern __shared__ float sm[];   // per warp: p[T] + ds[T] + k[hd] + v[hd]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kAttWarps + warp;
  if (row >= (int64_t)B * H * T) return;
  const int j = (int)(row % T);
  const int h = (int)((row / T) % H);
  const int b = (int)(row / ((int64_t)T * H));
  const int64_t tok_stride = 3LL * H * hd;
  const int64_t o_stride = (int64_t)H * hd;
  float* ps = sm + (int64_t)warp * (2 * T + 2 * kMaxHD);
  float* dss = ps + T;
  float* ks = dss + T;
  float* vs = ks + kMaxHD;
  const TA* kp = qkv + ((int64_t)b * T + j) * tok_stride + (int64_t)H * hd + h * hd;
  const TA* vp = kp + (int64_t)H * hd;
  for (int d = lane; d < hd; d += 32) {
    ks[d] = to_f(kp[d]) * scale;
    vs[d] = to_f(vp[d]);
  }
  __syncwarp();
  const TA* qbase = qkv + (int64_t)b * T * tok_stride + h * hd;
  const TA* dobase = d_o + (int64_t)b * T * o_stride + h * hd;
  const int64_t stat = ((int64_t)b * H + h) * T;
  for (int i = lane; i < T; i += 32) {
    const TA* qp = qbase + (int64_t)i * tok_stride;
    const TA* gp = dobase + (int64_t)i * o_stride;
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < hd; d += 4) {
      F4 qv = load4(qp + d), gv = load4(gp + d);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s += qv.v[c] * ks[d + c];
        dp += gv.v[c] * vs[d + c];
      }
    }
    float p = expf(s - lse[stat + i]);
    ps[i] = p;
    dss[i] = p * (dp - delta[stat + i]);
  }
  __syncwarp();
  float acck[kMaxHD / 32], accv[kMaxHD / 32];
#pragma unroll
  for (int c = 0; c < kMaxHD / 32; ++c) acck[c] = accv[c] = 0.f;
  for (int i = 0; i < T; ++i) {
    const float p = ps[i], w = dss[i];
    const TA* qp = qbase + (int64_t)i * tok_stride;
    const TA* gp = dobase + (int64_t)i * o_stride;
#pragma unroll
    for (int c = 0; c < kMaxHD / 32; ++c) {
      int d = lane + c * 32;
      if (d < hd) {
        accv[c] += p * to_f(gp[d]);
        acck[c] += w * to_f(qp[d]);
      }
    }
  }
  TA* dk = dqkv + ((int64_t)b * T + j) * tok_stride + (int64_t)H * hd + h * hd;
  TA* dv = dk + (int64_t)H * hd;
#pragma unroll
  for (int c = 0; c < kMaxHD / 32; ++c) {
    int d = lane + c * 32;
    if (d < hd) {
      dk[d] = from_f<TA>(acck[c] * scale);
      dv[d] = from_f<TA>(accv[c]);
    }
  }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) REED_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

int attn_simt_fwd(int act_dtype, const void* qkv, void* o, float* lse, int B, int T, int H, int hd, cudaStream_t st) {
  REED_REQUIRE(hd % 4 == 0 && hd <= kMaxHD, "attention head_dim must be a multiple of 4 and <= %d, got %d", kMaxHD, hd);
  const int64_t rows = (int64_t)B * H * T;
  if (rows == 0) return 0;
  const float scale = 1.f / sqrtf((float)hd);
  size_t smem = sizeof(float) * kAttWarps * (T + kMaxHD);
  REED_REQUIRE(smem <= 200 * 1024, "attention sequence too long for the SIMT kernel (T=%d)", T);
  dim3 grid(ceil_div(rows, kAttWarps));
  if (act_dtype == kBF16) {
    if (set_smem(attn_simt_fwd_kernel<bf16>, smem)) return 1;
    attn_simt_fwd_kernel<bf16><<<grid, kAttWarps * 32, smem, st>>>((const bf16*)qkv, (bf16*)o, lse, B, T, H, hd, scale);
  } else {
    if (set_smem(attn_simt_fwd_kernel<float>, smem)) return 1;
    attn_simt_fwd_kernel<float><<<grid, kAttWarps * 32, smem, st>>>((const float*)qkv, (float*)o, lse, B, T, H, hd, scale);
  }
  REED_LAUNCH_CHECK();
  return 0;
}

int attn_simt_bwd(int act_dtype, const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                  float* delta, int B, int T, int H, int hd, cudaStream_t st) {
  REED_REQUIRE(hd % 4 == 0 && hd <= kMaxHD, "attention head_dim must be a multiple of 4 and <= %d, got %d", kMaxHD, hd);
  const int64_t rows = (int64_t)B * H * T;
  if (rows == 0) return 0;
  const float scale = 1.f / sqrtf((float)hd);
  size_t smem_q = sizeof(float) * kAttWarps * (T + 2 * kMaxHD);
  size_t smem_kv = sizeof(float) * kAttWarps * (2 * T + 2 * kMaxHD);
  REED_REQUIRE(smem_kv <= 200 * 1024, "attention sequence too long for the SIMT kernel (T=%d)", T);
  dim3 grid(ceil_div(rows, kAttWarps));
  if (act_dtype == kBF16) {
    if (set_smem(attn_simt_bwd_q_kernel<bf16>, smem_q) || set_smem(attn_simt_bwd_kv_kernel<bf16>, smem_kv)) return 1;
    attn_simt_bwd_q_kernel<bf16><<<grid, kAttWarps * 32, smem_q, st>>>((const bf16*)qkv, (const bf16*)o, (const bf16*)d_o,
                                                                      lse, (bf16*)dqkv, delta, B, T, H, hd, scale);
    attn_simt_bwd_kv_kernel<bf16><<<grid, kAttWarps * 32, smem_kv, st>>>((const bf16*)qkv, (const bf16*)d_o, lse, delta,
                                                                        (bf16*)dqkv, B, T, H, hd, scale);
  } else {
    if (set_smem(attn_simt_bwd_q_kernel<float>, smem_q) || set_smem(attn_simt_bwd_kv_kernel<float>, smem_kv)) return 1;
    attn_simt_bwd_q_kernel<float><<<grid, kAttWarps * 32, smem_q, st>>>((const float*)qkv, (const float*)o,
                                                                       (const float*)d_o, lse, (float*)dqkv, delta, B, T,
                                                                       H, hd, scale);
    attn_simt_bwd_kv_kernel<float><<<grid, kAttWarps * 32, smem_kv, st>>>((const float*)qkv, (const float*)d_o, lse,
                                                                         delta, (float*)dqkv, B, T, H, hd, scale);
  }
  REED_LAUNCH_CHECK();
  return 0;
}

}  // namespace reed
