// Fused flash-style attention for the 256/1024-token patch sequence (bf16 operands, fp32 softmax/accumulators),
// forward and backward, reading Q/K/V straight out of the packed qkv GEMM output [B,T,3,H,hd] and writing the
// context as [B,T,H,hd] (the proj GEMM's A operand) and dqkv in the packed layout (the qkv dgrad/wgrad operand):
// no head-major transposes anywhere.  head_dim 64 (S/B/L) and 72 (XL; padded to 80 in shared memory only).
//
//   forward : one CTA = 64 query rows of one (batch, head); KV streamed in 64-row blocks with cp.async double
//             buffering; online softmax in the exp2 domain; S/P never leave registers.
//   backward: two atomic-free kernels - dQ (CTA = 64 query rows, loops over KV; also emits delta = rowsum(dO*O))
//             and dK/dV (CTA = 64 key rows, loops over queries, works on S^T so P^T/dS^T feed the MMAs directly).
//
// Tensor-core path: warp-level mma.sync m16n8k16 (bf16 -> fp32) with ldmatrix operand fetch.  Attention is 3.5 % of
// the SiT-XL/2 step FLOPs at T=256; the tcgen05/TMEM version is future work.
//
// Reference semantics: timm Attention.forward with fused_attn (F.scaled_dot_product_attention, scale hd^-0.5),
// imported at /root/reference/image/models/sit.py:13 and called at sit.py:134.
#include "common.cuh"

namespace reed {

constexpr int kBlk = 64;          // query rows per CTA == kv rows per block
constexpr int kAttnThreads = 128; // 4 warps x 16 rows

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int HD> struct AttnDims {
  static constexpr int DP = (HD + 15) / 16 * 16;   // contraction / output width in smem (72 -> 80)
  static constexpr int DS = DP + 8;                // smem row pitch (elements): ldmatrix rows land in distinct banks
  static constexpr int KSTEPS = DP / 16;
  static constexpr int NT = DP / 8;                // 8-wide output tiles over head_dim
  static constexpr int CHUNKS = HD / 8;            // 16-byte chunks per global row
  static constexpr int TILE = kBlk * DS;           // elements per 64-row smem tile
};

// copy a 64 x HD tile (rows row0..row0+63 of a [.., stride] matrix) into smem with cp.async
template <int HD>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int64_t stride, int tid) {
  using A = AttnDims<HD>;
  for (int i = tid; i < kBlk * A::CHUNKS; i += kAttnThreads) {
    int r = i / A::CHUNKS, c = i % A::CHUNKS;
    cp_async16(dst + r * A::DS + c * 8, src + (int64_t)r * stride + c * 8);
  }
}
// zero the padding columns HD..DP-1 of `tiles` consecutive tiles (never overwritten by load_tile)
template <int HD>
__device__ __forceinline__ void zero_pad(bf16* base, int tiles, int tid) {
  using A = AttnDims<HD>;
  if (A::DP == HD) return;
  for (int i = tid; i < tiles * kBlk; i += kAttnThreads) {
    bf16* p = base + (int64_t)i * A::DS + HD;
#pragma unroll
    for (int j = 0; j < A::DP - HD; ++j) p[j] = __float2bfloat16(0.f);
  }
}

// acc[16 x 64] (+)= X_w[16 x DP] . Y[64 x DP]^T ; X_w rows = this warp's 16 rows of tile X; Y read non-transposed
template <int HD>
__device__ __forceinline__ void mma_xyT(float (&acc)[8][4], const bf16* X, const bf16* Y, int warp, int lane) {
  using A = AttnDims<HD>;
#pragma unroll
  for (int ks = 0; ks < A::KSTEPS; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, X + (warp * 16 + (lane & 15)) * A::DS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, Y + (np * 16 + (lane & 7) + (lane >> 4) * 8) * A::DS + ks * 16 + ((lane >> 3) & 1) * 8);
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// out[16 x DP] += P[16 x 64] . Z[64 x DP] ; P given as accumulator-layout registers (converted to bf16 A fragments),
// Z row-major [k][n] read with ldmatrix.trans
template <int HD>
__device__ __forceinline__ void mma_pz(float (&out)[AttnDims<HD>::NT][4], const float (&p)[8][4], const bf16* Z,
                                       int lane) {
  using A = AttnDims<HD>;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < A::NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, Z + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * A::DS + np * 16 + (lane >> 4) * 8);
      mma_bf16(out[2 * np], a, b[0], b[1]);
      mma_bf16(out[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o,
                                                                 float* __restrict__ lse, int T, int H, float scale_log2) {
  using A = AttnDims<HD>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + A::TILE;          // 2 stages
  bf16* sV = sK + 2 * A::TILE;      // 2 stages
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int64_t tok = 3LL * H * HD;
  const bf16* qbase = qkv + ((int64_t)b * T + qb * kBlk) * tok + h * HD;
  const bf16* kbase = qkv + (int64_t)b * T * tok + (int64_t)H * HD + h * HD;
  const bf16* vbase = kbase + (int64_t)H * HD;

  zero_pad<HD>(sQ, 5, tid);
  load_tile<HD>(sQ, qbase, tok, tid);
  load_tile<HD>(sK, kbase, tok, tid);
  load_tile<HD>(sV, vbase, tok, tid);
  cp_async_commit();

  float oacc[A::NT][4];
#pragma unroll
  for (int i = 0; i < A::NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  const int nblk = T / kBlk;
  for (int kb = 0; kb < nblk; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < nblk) {
      load_tile<HD>(sK + (st ^ 1) * A::TILE, kbase + (int64_t)(kb + 1) * kBlk * tok, tok, tid);
      load_tile<HD>(sV + (st ^ 1) * A::TILE, vbase + (int64_t)(kb + 1) * kBlk * tok, tok, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    mma_xyT<HD>(s, sQ, sK + st * A::TILE, warp, lane);

    // online softmax (rows g = lane/4 and g+8 of this warp's 16), exp2 domain
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fmaxf(s[i][2 * r], s[i][2 * r + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[r], mx * scale_log2);
      const float corr = exp2f(m_run[r] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float p0 = exp2f(s[i][2 * r] * scale_log2 - m_new);
        float p1 = exp2f(s[i][2 * r + 1] * scale_log2 - m_new);
        s[i][2 * r] = p0;
        s[i][2 * r + 1] = p1;
        sum += p0 + p1;
      }
      l_run[r] = l_run[r] * corr + sum;
      m_run[r] = m_new;
#pragma unroll
      for (int i = 0; i < A::NT; ++i) {
        oacc[i][2 * r] *= corr;
        oacc[i][2 * r + 1] *= corr;
      }
    }
    mma_pz<HD>(oacc, s, sV + st * A::TILE, lane);
    __syncthreads();   // everyone done with stage st before it is refilled two iterations later
  }

  // finalise: O / l, lse (natural log), store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.f / l;
    const int row = qb * kBlk + warp * 16 + (lane >> 2) + 8 * r;
    if ((lane & 3) == 0) lse[((int64_t)b * H + h) * T + row] = (m_run[r] + log2f(l)) * 0.6931471805599453f;
    bf16* orow = o + ((int64_t)b * T + row) * ((int64_t)H * HD) + h * HD;
#pragma unroll
    for (int i = 0; i < A::NT; ++i) {
      int col = i * 8 + (lane & 3) * 2;
      if (col < HD)
        *reinterpret_cast<__nv_bfloat162*>(orow + col) = __floats2bfloat162_rn(oacc[i][2 * r] * inv, oacc[i][2 * r + 1] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// backward, part 1: dQ (and delta).  CTA = 64 query rows; loops over KV blocks.
// ------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_dq_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
    const float* __restrict__ lse, bf16* __restrict__ dqkv, float* __restrict__ delta, int T, int H, float scale,
    float scale_log2) {
  using A = AttnDims<HD>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sDO = sQ + A::TILE;
  bf16* sK = sDO + A::TILE;         // 2 stages
  bf16* sV = sK + 2 * A::TILE;      // 2 stages
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int64_t tok = 3LL * H * HD, otok = (int64_t)H * HD;
  const bf16* qbase = qkv + ((int64_t)b * T + qb * kBlk) * tok + h * HD;
  const bf16* kbase = qkv + (int64_t)b * T * tok + (int64_t)H * HD + h * HD;
  const bf16* vbase = kbase + (int64_t)H * HD;
  const bf16* dobase = d_o + ((int64_t)b * T + qb * kBlk) * otok + h * HD;
  const bf16* obase = o + ((int64_t)b * T + qb * kBlk) * otok + h * HD;

  zero_pad<HD>(sQ, 6, tid);
  load_tile<HD>(sQ, qbase, tok, tid);
  load_tile<HD>(sDO, dobase, otok, tid);
  load_tile<HD>(sK, kbase, tok, tid);
  load_tile<HD>(sV, vbase, tok, tid);
  cp_async_commit();

  // delta for this thread's two rows, straight from global (each quad covers a row)
  float dl[2], ls[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = warp * 16 + (lane >> 2) + 8 * r;
    const bf16* op = obase + (int64_t)row * otok;
    const bf16* gp = dobase + (int64_t)row * otok;
    float acc = 0.f;
    for (int c = (lane & 3) * 2; c < HD; c += 8) {
      __nv_bfloat162 ov = *reinterpret_cast<const __nv_bfloat162*>(op + c);
      __nv_bfloat162 gv = *reinterpret_cast<const __nv_bfloat162*>(gp + c);
      acc += __low2float(ov) * __low2float(gv) + __high2float(ov) * __high2float(gv);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    dl[r] = acc;
    const int64_t sidx = ((int64_t)b * H + h) * T + qb * kBlk + row;
    if ((lane & 3) == 0) delta[sidx] = acc;
    ls[r] = lse[sidx] * 1.4426950408889634f;   // to the exp2 domain
  }

  float dq[A::NT][4];
#pragma unroll
  for (int i = 0; i < A::NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;

  const int nblk = T / kBlk;
  for (int kb = 0; kb < nblk; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < nblk) {
      load_tile<HD>(sK + (st ^ 1) * A::TILE, kbase + (int64_t)(kb + 1) * kBlk * tok, tok, tid);
      load_tile<HD>(sV + (st ^ 1) * A::TILE, vbase + (int64_t)(kb + 1) * kBlk * tok, tok, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
    mma_xyT<HD>(s, sQ, sK + st * A::TILE, warp, lane);     // S  = Q K^T
    mma_xyT<HD>(dp, sDO, sV + st * A::TILE, warp, lane);   // dP = dO V^T
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = j >> 1;
        const float p = exp2f(s[i][j] * scale_log2 - ls[r]);
        s[i][j] = p * (dp[i][j] - dl[r]);                  // dS (without the softmax scale)
      }
    mma_pz<HD>(dq, s, sK + st * A::TILE, lane);            // dQ += dS K
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = qb * kBlk + warp * 16 + (lane >> 2) + 8 * r;
    bf16* out = dqkv + ((int64_t)b * T + row) * tok + h * HD;
#pragma unroll
    for (int i = 0; i < A::NT; ++i) {
      int col = i * 8 + (lane & 3) * 2;
      if (col < HD)
        *reinterpret_cast<__nv_bfloat162*>(out + col) = __floats2bfloat162_rn(dq[i][2 * r] * scale, dq[i][2 * r + 1] * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// backward, part 2: dK and dV.  CTA = 64 key rows; loops over query blocks; everything is computed transposed.
// ------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_dkv_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, const float* __restrict__ lse,
    const float* __restrict__ delta, bf16* __restrict__ dqkv, int T, int H, float scale, float scale_log2) {
  using A = AttnDims<HD>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);
  bf16* sV = sK + A::TILE;
  bf16* sQ = sV + A::TILE;          // 2 stages
  bf16* sDO = sQ + 2 * A::TILE;     // 2 stages
  float* sL = reinterpret_cast<float*>(sDO + 2 * A::TILE);   // [2][64] lse (exp2 domain)
  float* sD = sL + 2 * kBlk;                                  // [2][64] delta
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int64_t tok = 3LL * H * HD, otok = (int64_t)H * HD;
  const bf16* qbase = qkv + (int64_t)b * T * tok + h * HD;
  const bf16* kbase = qbase + (int64_t)H * HD + (int64_t)kb * kBlk * tok;
  const bf16* vbase = kbase + (int64_t)H * HD;
  const bf16* dobase = d_o + (int64_t)b * T * otok + h * HD;
  const float* lbase = lse + ((int64_t)b * H + h) * T;
  const float* dbase = delta + ((int64_t)b * H + h) * T;

  zero_pad<HD>(sK, 6, tid);
  load_tile<HD>(sK, kbase, tok, tid);
  load_tile<HD>(sV, vbase, tok, tid);
  load_tile<HD>(sQ, qbase, tok, tid);
  load_tile<HD>(sDO, dobase, otok, tid);
  cp_async_commit();
  if (tid < kBlk) {
    sL[tid] = lbase[tid] * 1.4426950408889634f;
    sD[tid] = dbase[tid];
  }

  float dk[A::NT][4], dv[A::NT][4];
#pragma unroll
  for (int i = 0; i < A::NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;

  const int nblk = T / kBlk;
  for (int qb = 0; qb < nblk; ++qb) {
    const int st = qb & 1;
    if (qb + 1 < nblk) {
      load_tile<HD>(sQ + (st ^ 1) * A::TILE, qbase + (int64_t)(qb + 1) * kBlk * tok, tok, tid);
      load_tile<HD>(sDO + (st ^ 1) * A::TILE, dobase + (int64_t)(qb + 1) * kBlk * otok, otok, tid);
      cp_async_commit();
      if (tid < kBlk) {
        sL[(st ^ 1) * kBlk + tid] = lbase[(qb + 1) * kBlk + tid] * 1.4426950408889634f;
        sD[(st ^ 1) * kBlk + tid] = dbase[(qb + 1) * kBlk + tid];
      }
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
    mma_xyT<HD>(s, sK, sQ + st * A::TILE, warp, lane);      // S^T  = K Q^T   (rows: this warp's keys, cols: queries)
    mma_xyT<HD>(dp, sV, sDO + st * A::TILE, warp, lane);    // dP^T = V dO^T
    const float* L = sL + st * kBlk;
    const float* Dl = sD + st * kBlk;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = i * 8 + (lane & 3) * 2 + (j & 1);
        const float p = exp2f(s[i][j] * scale_log2 - L[q]);
        s[i][j] = p;                                        // P^T
        dp[i][j] = p * (dp[i][j] - Dl[q]);                  // dS^T
      }
    mma_pz<HD>(dv, s, sDO + st * A::TILE, lane);            // dV += P^T dO
    mma_pz<HD>(dk, dp, sQ + st * A::TILE, lane);            // dK += dS^T Q
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = kb * kBlk + warp * 16 + (lane >> 2) + 8 * r;
    bf16* outk = dqkv + ((int64_t)b * T + row) * tok + (int64_t)H * HD + h * HD;
    bf16* outv = outk + (int64_t)H * HD;
#pragma unroll
    for (int i = 0; i < A::NT; ++i) {
      int col = i * 8 + (lane & 3) * 2;
      if (col < HD) {
        *reinterpret_cast<__nv_bfloat162*>(outk + col) = __floats2bfloat162_rn(dk[i][2 * r] * scale, dk[i][2 * r + 1] * scale);
        *reinterpret_cast<__nv_bfloat162*>(outv + col) = __floats2bfloat162_rn(dv[i][2 * r], dv[i][2 * r + 1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------
bool attn_mma_supported(int T, int hd) { return T >= kBlk && T % kBlk == 0 && (hd == 64 || hd == 72); }

template <int HD>
static int fwd_launch(const void* qkv, void* o, float* lse, int B, int T, int H, cudaStream_t st) {
  using A = AttnDims<HD>;
  const int smem = 5 * A::TILE * 2;
  static bool done = false;
  if (!done) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  attn_fwd_kernel<HD><<<dim3(T / kBlk, H, B), kAttnThreads, smem, st>>>((const bf16*)qkv, (bf16*)o, lse, T, H, scale_log2);
  REED_LAUNCH_CHECK();
  return 0;
}

template <int HD>
static int bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta,
                      int B, int T, int H, cudaStream_t st) {
  using A = AttnDims<HD>;
  const int smem_q = 6 * A::TILE * 2;
  const int smem_kv = 6 * A::TILE * 2 + 4 * kBlk * 4;
  static bool done = false;
  if (!done) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    done = true;
  }
  const float scale = 1.f / sqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(T / kBlk, H, B);
  attn_bwd_dq_kernel<HD><<<grid, kAttnThreads, smem_q, st>>>((const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse,
                                                            (bf16*)dqkv, delta, T, H, scale, scale_log2);
  attn_bwd_dkv_kernel<HD><<<grid, kAttnThreads, smem_kv, st>>>((const bf16*)qkv, (const bf16*)d_o, lse, delta,
                                                              (bf16*)dqkv, T, H, scale, scale_log2);
  REED_LAUNCH_CHECK();
  return 0;
}

int attn_mma_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int hd, cudaStream_t st) {
  if (B == 0) return 0;
  if (hd == 64) return fwd_launch<64>(qkv, o, lse, B, T, H, st);
  if (hd == 72) return fwd_launch<72>(qkv, o, lse, B, T, H, st);
  return fail("tensor-core attention: head_dim %d unsupported", hd);
}

int attn_mma_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta, int B,
                 int T, int H, int hd, cudaStream_t st) {
  if (B == 0) return 0;
  if (hd == 64) return bwd_launch<64>(qkv, o, d_o, lse, dqkv, delta, B, T, H, st);
  if (hd == 72) return bwd_launch<72>(qkv, o, d_o, lse, dqkv, delta, B, T, H, st);
  return fail("tensor-core attention: head_dim %d unsupported", hd);
}

}  // namespace reed
