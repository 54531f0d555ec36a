// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands staged in
// shared memory by TMA with the 128-byte swizzle, persistent over output tiles, warp-specialised:
//   warp 0   : TMA producer          (one elected lane)
//   warp 1   : tcgen05.mma issuer    (one elected lane), commits free smem stages / publish accumulators
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue - tcgen05.ld the 128 x BN accumulator, transpose through smem so global accesses are
//              row-coalesced, apply the fused epilogue (bias / GELU / SiLU / gate*y+residual / act') and store.
// Two accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
//
//   D[M,N] = epi( A[M,K] . B[N,K]^T )        A, B bf16; each either K-major or MN-major in global memory:
//   forward  y  = x W^T      : A = x  (K-major),  B = W  (K-major)
//   dgrad    dx = dy W       : A = dy (K-major),  B = W  (MN-major: stored [N_contract, K_out])
//   wgrad    dW = dy^T x     : A = dy (MN-major), B = x  (MN-major)
// This covers the qkv / proj / fc1 / fc2 / adaLN / projector linears of /root/reference/image/models/sit.py
// (timm Attention.qkv/proj, Mlp.fc1/fc2 at sit.py:114-124; adaLN 125-128; build_mlp 17-24) and their backward.
#include <cuda.h>
#include "common.cuh"

namespace reed {

constexpr int BM = 128;         // UMMA M (cta_group::1)
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 256;
constexpr int kEpiWarps = 4;
constexpr int kStageCols = 32;  // accumulator columns moved per tcgen05.ld
constexpr int kStagePitch = 36; // floats; 144 B row pitch keeps float4 smem accesses conflict-free

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  uint64_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1ull << 26)) {   // ~seconds: a protocol bug must surface as an error, never as a hung GPU
      printf("reed gemm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1.
//   K-major : rows of 128 B, 8-row swizzle atoms 1024 B apart            -> SBO = 1024, LBO unused
//   MN-major: each TMA box is 64 (mn) x BK (k) elements: k-rows of 128 B, -> SBO = 1024 (next 8 k-rows),
//             the next 64 mn-elements live in the next box               -> LBO = BK * 128 B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

// cute::UMMA::InstrDescriptor: D fp32, A/B bf16, dense
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 32 * kStagePitch * 4;
  static constexpr int kBudget = 227 * 1024 - 1024 /*align slack*/ - kStagingBytes - 256 /*barriers*/;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStagingBytes + 256;
  static constexpr int kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
};

template <int BN, int A_MN, int B_MN, typename TD>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    TD* __restrict__ D, int64_t ldd, int M, int N, int K, EpiParams ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* staging = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full = bars;              // [S]
  uint64_t* empty = bars + S;         // [S]
  uint64_t* tfull = bars + 2 * S;     // [2]
  uint64_t* tempty = bars + 2 * S + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ============================== TMA producer ==============================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        mbar_expect_tx(&full[stage], Cfg::kStageBytes);
        const int k0 = kb * BK;
        if (A_MN) {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_2d(&map_a, &full[stage], sa + c * (BK * 128), m0 + c * 64, k0);
        } else {
          tma_load_2d(&map_a, &full[stage], sa, k0, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(&map_b, &full[stage], sb + c * (BK * 128), n0 + c * 64, k0);
        } else {
          tma_load_2d(&map_b, &full[stage], sb, k0, n0);
        }
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: +32 B inside the swizzled 128 B row per 16-element k step; MN-major: +16 k-rows of 128 B
          const uint64_t da = A_MN ? make_smem_desc(sa + k * (UMMA_K * 128), BK * 128, 1024)
                                   : make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc(sb + k * (UMMA_K * 128), BK * 128, 1024)
                                   : make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
          umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    float* st = staging + q * 32 * kStagePitch;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      const int row_base = m0 + q * 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += kStageCols) {
        if (n0 + c0 >= N) break;                 // warp-uniform
        float v[32];
        tmem_ld32(taddr + c0, v);
        // thread = accumulator row: park the 32 columns in smem ...
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + lane * kStagePitch + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        // ... and pick them up row-coalesced: 8 lanes cover one 32-column row segment, 4 rows per instruction
        const int cc = (lane & 7) * 4;
        const int col = n0 + c0 + cc;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int row = row_base + r;
          float4 f = *reinterpret_cast<const float4*>(st + r * kStagePitch + cc);
          if (row < M && col < N) epilogue_store4<TD, bf16>(ep, D, ldd, row, col, F4{{f.x, f.y, f.z, f.w}});
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row pitch ld (elements); box = box_rows x 64 cols
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  REED_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REED_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r,
               (long long)rows, (long long)cols, (long long)ld, box_rows);
  return 0;
}

template <int BN, int A_MN, int B_MN, typename TD>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd, int M, int N, int K,
                  const EpiParams& ep, cudaStream_t st, int max_ctas) {
  using Cfg = GemmCfg<BN>;
  static_assert(Cfg::kStages >= 3, "pipeline too shallow");
  auto kernel = gemm_tcgen05_kernel<BN, A_MN, B_MN, TD>;
  static bool configured = false;   // per template instance
  if (!configured) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  int grid = tiles < max_ctas ? tiles : max_ctas;
  kernel<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(ma, mb, (TD*)D, ldd, M, N, K, ep);
  REED_LAUNCH_CHECK();
  return 0;
}

template <int BN, typename TD>
static int launch_major(int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd, int M,
                        int N, int K, const EpiParams& ep, cudaStream_t st, int max_ctas) {
  if (!a_mn && !b_mn) return launch<BN, 0, 0, TD>(ma, mb, D, ldd, M, N, K, ep, st, max_ctas);
  if (!a_mn && b_mn) return launch<BN, 0, 1, TD>(ma, mb, D, ldd, M, N, K, ep, st, max_ctas);
  if (a_mn && b_mn) return launch<BN, 1, 1, TD>(ma, mb, D, ldd, M, N, K, ep, st, max_ctas);
  return launch<BN, 1, 0, TD>(ma, mb, D, ldd, M, N, K, ep, st, max_ctas);
}

bool gemm_tcgen05_supported(int64_t lda, int64_t ldb, int64_t ldd, const void* A, const void* B, int M, int N, int K) {
  return lda % 8 == 0 && ldb % 8 == 0 && ldd % 4 == 0 && N % 8 == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0 &&
         M >= 1 && N >= 64 && K >= 64;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
  }
  return n;
}

int gemm_tcgen05(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D, int64_t ldd,
                 int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  // tile width: the largest of 256/192/128 that divides N (fewest wasted MMA columns), else by padding waste
  int bn = 128;
  if (N % 256 == 0) bn = 256;
  else if (N % 192 == 0) bn = 192;
  else if (N % 128 == 0) bn = 128;
  else {
    int best_waste = 1 << 30;
    for (int cand : {256, 192, 128}) {
      int waste = ceil_div(N, cand) * cand - N;
      if (waste < best_waste) { best_waste = waste; bn = cand; }
    }
  }
  // prefer a narrower tile when the wide one cannot fill the machine
  const int ctas = sm_count();
  if (bn > 128 && ceil_div(M, BM) * ceil_div(N, bn) < ctas) bn = 128;

  CUtensorMap ma, mb;
  // K-major operand [MN, K] row-major: box = (BM|BN) rows x 64 k.  MN-major operand stored [K, MN]: box = 64 k-rows x 64 mn.
  if (a_mn) { if (make_map(&ma, A, K, M, lda, BK)) return 1; } else { if (make_map(&ma, A, M, K, lda, BM)) return 1; }
  if (b_mn) { if (make_map(&mb, B, K, N, ldb, BK)) return 1; } else { if (make_map(&mb, B, N, K, ldb, bn)) return 1; }

#define GO(BNV)                                                                                              \
  (d_dtype == kF32 ? launch_major<BNV, float>(a_mn, b_mn, ma, mb, D, ldd, M, N, K, ep, st, ctas)             \
                   : launch_major<BNV, bf16>(a_mn, b_mn, ma, mb, D, ldd, M, N, K, ep, st, ctas))
  if (bn == 256) return GO(256);
  if (bn == 192) return GO(192);
  return GO(128);
#undef GO
}

}  // namespace reed
