// fp32-accurate SIMT GEMM with the shared epilogues.  This is the compute path of the fp32 precision mode
// (parity bar 1e-5 relative; tcgen05 has no true-fp32 MMA) and of shapes the tcgen05 kernel does not take
// (dimensions that are not multiples of 8).  D[M,N] = epi(A[M,K] . B[N,K]^T), generic element strides.
#include "common.cuh"

namespace reed {

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename TA, typename TD>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, int64_t a_sm, int64_t a_sk,
                                                         const TA* __restrict__ B, int64_t b_sn, int64_t b_sk,
                                                         TD* __restrict__ D, int64_t ldd, int M, int N, int K,
                                                         EpiParams ep, int k_per_split) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int tx = tid & 15, ty = tid >> 4;     // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: choose the fastest-varying thread index along the contiguous dimension of each operand
  const bool a_k_contig = a_sk == 1, b_k_contig = b_sk == 1;
  // split-K (gridDim.z > 1): each z-slice reduces its own k-range and atomically adds into a zeroed fp32 D
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  for (int k0 = k_begin; k0 < k_end; k0 += SBK) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int e = tid + it * 256;             // 1024 elements per tile
      int mm, kk;
      if (a_k_contig) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < k_end) ? to_f(A[gm * a_sm + gk * a_sk]) : 0.f;
      int nn;
      if (b_k_contig) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
      int gn = n0 + nn;
      gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < k_end) ? to_f(B[gn * b_sn + gk * b_sk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int col = n0 + tx * 4;
  if (col >= N) return;   // N % 4 == 0 is required by the host wrapper
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    if (gridDim.z > 1) {
      if constexpr (sizeof(TD) == 4) {
        float* d = reinterpret_cast<float*>(D) + (int64_t)row * ldd + col;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][j];
          if (blockIdx.z == 0 && ep.bias != nullptr) v += ep.bias[col + j];
          atomicAdd(d + j, v);
        }
      }
    } else {
      epilogue_store4<TD, TA>(ep, D, ldd, row, col, F4{{acc[i][0], acc[i][1], acc[i][2], acc[i][3]}});
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient over a handful of rows: D[N, K] (+)= sum_b dy[b, N] * x[b, K], b <= 64 (the adaLN / embedder linears,
// whose contraction is the batch).  An outer-product stream: 2 small operand tiles in shared memory, every thread
// owns an 8 x 4 output block, rows written as coalesced float4 - bounded by the HBM write of N*K floats.
// dy and x may differ in type (the adaLN modulation gradient stays fp32, silu(c) is bf16); with `db` the CTAs of the
// first column tile also add the column sums of dy into it (the bias gradient, no separate pass over dy).
// ------------------------------------------------------------------------------------------------
constexpr int kOwN = 64, kOwK = 128;

// rows [0, B) x `width` columns starting at column c0 of a row-major matrix -> fp32 shared memory [B][width]
template <typename T, int WIDTH>
__device__ __forceinline__ void ow_stage(const T* __restrict__ src, int64_t ld, int c0, int cols, int B, float* dst, int tid) {
  constexpr int kVec = sizeof(T) == 2 ? 8 : 4;     // elements per 16-byte load
  const bool fast = c0 + WIDTH <= cols && ld % kVec == 0 && ((uintptr_t)src & 15) == 0;
  if (fast) {
    for (int e = tid; e < B * (WIDTH / kVec); e += 256) {
      const int b = e / (WIDTH / kVec), c = e % (WIDTH / kVec);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)b * ld + c0 + c * kVec));
      float* o = dst + b * WIDTH + c * kVec;
      if constexpr (sizeof(T) == 2) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
        *reinterpret_cast<float4*>(o) = make_float4(__low2float(h[0]), __high2float(h[0]), __low2float(h[1]), __high2float(h[1]));
        *reinterpret_cast<float4*>(o + 4) = make_float4(__low2float(h[2]), __high2float(h[2]), __low2float(h[3]), __high2float(h[3]));
      } else {
        *reinterpret_cast<uint4*>(o) = v;
      }
    }
  } else {
    for (int e = tid; e < B * WIDTH; e += 256) {
      const int b = e / WIDTH, c = e % WIDTH;
      dst[b * WIDTH + c] = (c0 + c < cols) ? to_f(src[(int64_t)b * ld + c0 + c]) : 0.f;
    }
  }
}

// CTA = 64 x 128 outputs, thread = 8 rows x 4 columns; the batch rows of dy / x are staged in shared memory as fp32
// (dynamic, B * 192 floats); a CTA may walk several column tiles of its row strip (kt_per_cta), the launcher uses one.
// Row pairs ride on the packed fp32 FMA (FFMA2): 16 issue slots per batch row for 32 FMAs, operands from three
// LDS.128 (the dy values are a warp-wide broadcast).  The kernel is bound by its M*N fp32 writes.
template <typename TDY, typename TX>
__global__ void __launch_bounds__(256) outer_wgrad_kernel(const TDY* __restrict__ dy, int64_t ld_dy, const TX* __restrict__ x,
                                                           int64_t ld_x, float* __restrict__ D, int64_t ldd, int N, int K,
                                                           int B, int accumulate, float* __restrict__ db, int kt_per_cta) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  extern __shared__ __align__(16) float ow_smem[];
  float* sdy = ow_smem;                    // [B][kOwN]
  float* sx = ow_smem + B * kOwN;          // [B][kOwK]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * kOwN;
  const int tiles_k = (K + kOwK - 1) / kOwK;
  const int kt0 = blockIdx.x * kt_per_cta, kt1 = min(kt0 + kt_per_cta, tiles_k);
  pdl_wait();
  ow_stage<TDY, kOwN>(dy, ld_dy, n0, N, B, sdy, tid);
  const int tk = (tid & 31) * 4, tn = (tid >> 5) * 8;
  for (int kt = kt0; kt < kt1; ++kt) {       // the CTA's column tiles share the staged dy tile
    const int k0 = kt * kOwK;
    if (kt > kt0) __syncthreads();           // everyone is done reading the previous x tile
    ow_stage<TX, kOwK>(x, ld_x, k0, K, B, sx, tid);
    __syncthreads();
    if (db != nullptr && kt == 0 && tid < kOwN && n0 + tid < N) {     // bias gradient: column sums of dy
      float sum = 0.f;
      for (int b = 0; b < B; ++b) sum += sdy[b * kOwN + tid];
      db[n0 + tid] += sum;
    }
    float2 acc[4][4];   // [row pair][column]: (row 2p, row 2p+1)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[p][j] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int b = 0; b < B; ++b) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sdy[b * kOwN + tn]);     // same address across the warp: broadcast
      const float4 a1 = *reinterpret_cast<const float4*>(&sdy[b * kOwN + tn + 4]);
      const float4 x4 = *reinterpret_cast<const float4*>(&sx[b * kOwK + tk]);
      const float2 ap[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
      const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xx = make_float2(xv[j], xv[j]);
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p][j] = __ffma2_rn(ap[p], xx, acc[p][j]);
      }
    }
    const int col = k0 + tk;
    if (col >= K) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = n0 + tn + i;
      if (row >= N) continue;
      float* d = D + (int64_t)row * ldd + col;
      const int p = i >> 1;
      F4 o;
#pragma unroll
      for (int j = 0; j < 4; ++j) o.v[j] = (i & 1) ? acc[p][j].y : acc[p][j].x;
      if (accumulate) {
        const F4 old = load4(d);
#pragma unroll
        for (int j = 0; j < 4; ++j) o.v[j] += old.v[j];
      }
      store4(d, o);
    }
  }
}

// D[N = columns of dy, K = columns of x] (+)= dy^T x, db[N] += column sums of dy (optional); B <= 64 rows
template <typename TDY, typename TX>
static int outer_wgrad_launch(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, float* D, int64_t ldd, int N, int K,
                              int B, int accumulate, float* db, cudaStream_t st) {
  auto kernel = outer_wgrad_kernel<TDY, TX>;
  const int smem = B * (kOwN + kOwK) * 4;
  static bool configured = false;
  if (!configured) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * (kOwN + kOwK) * 4));
    configured = true;
  }
  const int tiles_k = ceil_div(K, kOwK), tiles_n = ceil_div(N, kOwN);
  // one column tile per CTA: three CTAs fit an SM (70 registers) and overlap each other's load / compute / store phases;
  // CTAs that walk three tiles each (one wave of 324) measured 20.7 us against 15.5 us for 972 single-tile CTAs
  const int kt_per_cta = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ceil_div(tiles_k, kt_per_cta), tiles_n);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  REED_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, (const TDY*)dy, ld_dy, (const TX*)x, ld_x, D, ldd, N, K, B, accumulate, db,
                                     kt_per_cta));
  return 0;
}

int outer_wgrad(int dy_dtype, const void* dy, int64_t ld_dy, int x_dtype, const void* x, int64_t ld_x, float* D, int64_t ldd,
                int N, int K, int B, int accumulate, float* db, cudaStream_t st) {
  REED_REQUIRE(B >= 1 && B <= 64, "outer_wgrad: %d rows (1..64)", B);
  REED_REQUIRE(K % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)D & 15) == 0, "outer_wgrad: the output needs 16-byte rows");
  if (dy_dtype == kBF16 && x_dtype == kBF16) return outer_wgrad_launch<bf16, bf16>(dy, ld_dy, x, ld_x, D, ldd, N, K, B, accumulate, db, st);
  if (dy_dtype == kF32 && x_dtype == kBF16) return outer_wgrad_launch<float, bf16>(dy, ld_dy, x, ld_x, D, ldd, N, K, B, accumulate, db, st);
  if (dy_dtype == kF32 && x_dtype == kF32) return outer_wgrad_launch<float, float>(dy, ld_dy, x, ld_x, D, ldd, N, K, B, accumulate, db, st);
  return fail("outer_wgrad: unsupported operand types (dy %d, x %d)", dy_dtype, x_dtype);
}

int gemm_simt(int act_dtype, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D,
              int64_t ldd, int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  REED_REQUIRE(N % 4 == 0 && ldd % 4 == 0, "gemm_simt needs N %% 4 == 0 and ldd %% 4 == 0 (N=%d ldd=%lld)", N, (long long)ldd);
  // batch-contraction weight gradient (A = dy stored [K, M], B = x stored [K, N], K <= 64 rows): outer-product stream
  if (a_mn && b_mn && K <= 64 && d_dtype == kF32 && ep.kind == kEpiNone && ep.bias == nullptr && (int64_t)M * N >= 65536)
    return outer_wgrad(act_dtype, A, lda, act_dtype, B, ldb, (float*)D, ldd, M, N, K, ep.accumulate, nullptr, st);
  dim3 grid(ceil_div(N, SBN), ceil_div(M, SBM));
  int k_per_split = K > 0 ? K : 1;
  // few output tiles and a long reduction (patch-embed / final-layer wgrad): split K across the machine
  const int tiles = grid.x * grid.y;
  if (ep.kind == kEpiNone && d_dtype == kF32 && tiles * 2 <= kNumSMs && K >= 2048) {
    int splits = (2 * kNumSMs) / tiles;
    int max_splits = K / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits > 1) {
      k_per_split = ceil_div(ceil_div(K, splits), SBK) * SBK;
      grid.z = ceil_div(K, k_per_split);
      if (!ep.accumulate)
        REED_CHECK_CUDA(cudaMemset2DAsync(D, (size_t)ldd * 4, 0, (size_t)N * 4, (size_t)M, st));
    }
  }
  int64_t a_sm = a_mn ? 1 : lda, a_sk = a_mn ? lda : 1;
  int64_t b_sn = b_mn ? 1 : ldb, b_sk = b_mn ? ldb : 1;
#define GS(TA, TD) gemm_simt_kernel<TA, TD><<<grid, 256, 0, st>>>((const TA*)A, a_sm, a_sk, (const TA*)B, b_sn, b_sk, (TD*)D, ldd, M, N, K, ep, k_per_split)
  if (act_dtype == kF32 && d_dtype == kF32) GS(float, float);
  else if (act_dtype == kBF16 && d_dtype == kBF16) GS(bf16, bf16);
  else if (act_dtype == kBF16 && d_dtype == kF32) GS(bf16, float);
  else return fail("gemm_simt: fp32 activations with bf16 output are not a supported combination");
#undef GS
  REED_LAUNCH_CHECK();
  return 0;
}

}  // namespace reed
