// Host side of the tcgen05 GEMM: tensor maps, tile / cluster / split-K plan, dispatch to the cta_group::1 / ::2 kernels.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace reed {

struct EpiMaps { CUtensorMap d, o2, aux; };   // same layout as in gemm_tcgen05.cuh
constexpr int kMaxGroups = 32;
struct GroupMaps { CUtensorMap a[kMaxGroups], b[kMaxGroups]; int mode, per_group; };   // same layout as in gemm_tcgen05.cuh
int gemm_tc_launch_grouped(int bn, int b_mn, const GroupMaps& gm, void* D, int64_t ldd, int M, int N, int K, const EpiParams& ep,
                           cudaStream_t st, int grid, int stream_k);

int gemm_tc_launch_cg1(int bn, int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd,
                       int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st, int grid, int stream_k,
                       const EpiMaps* em);
int gemm_tc_launch_cg2(int bn, int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd,
                       int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st, int grid, int stream_k,
                       const EpiMaps* em);

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

// [32 rows x 32 columns] box over a row-major [rows, cols] matrix for the TMA epilogue: bf16 -> 64-byte rows with
// SWIZZLE_64B, fp32 -> 128-byte rows with SWIZZLE_128B (one epilogue warp's share of a 32-column chunk)
static int make_epi_map(CUtensorMap* map, const void* ptr, int dtype, int64_t rows, int64_t cols, int64_t ld, int box_cols = 32) {
  EncodeTiledFn enc = get_encode_fn();
  REED_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const int esz = dtype == kF32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  const int row_bytes = box_cols * esz;   // 128 / 64 / 32-byte box rows take the swizzle of the same span
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(map, dtype == kF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REED_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (epilogue box) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
               (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

static bool tma_ok_ptr(const void* p, int64_t ld, int esz) { return ((uintptr_t)p & 15) == 0 && (ld * esz) % 16 == 0; }

bool gemm_tcgen05_supported(int64_t lda, int64_t ldb, int64_t ldd, const void* A, const void* B, int M, int N, int K) {
  return lda % 8 == 0 && ldb % 8 == 0 && ldd % 4 == 0 && N % 8 == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0 &&
         M >= 1 && N >= 16 && K >= 16;   // K < 64: TMA zero-fills the rest of the 64-wide k-block
}

// SMs the persistent GEMM grids leave free (reed_gemm_reserve_sms): under data parallelism the NCCL all-reduce
// kernels of finished buckets run beside the backward GEMMs; a 148-CTA persistent grid that finds some SMs taken
// runs its last CTAs as a second wave and nearly doubles in time.
static int g_reserve_sms = -1;
void gemm_tcgen05_reserve_sms(int n) { g_reserve_sms = n < 0 ? 0 : n; }

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
  }
  if (g_reserve_sms < 0) g_reserve_sms = getenv("REED_GEMM_RESERVE_SMS") ? atoi(getenv("REED_GEMM_RESERVE_SMS")) : 0;
  int use = n - g_reserve_sms;
  use -= use & 1;
  return use < 2 ? 2 : use;
}

// Plan: cta_group (1 / 2), tile width BN, data-parallel or split (k slices for the last partial round).  Cost model per k-block and per CTA, in SM
// cycles: the MMA needs 2*BN (UMMA 128*CG x BN x 16 retires in BN/2 cycles), the operand fetch needs
// (16 KB of A + BN/CG * 128 B of B) / ~42 B/clk (the measured L2->SM share of one SM, B300_MICROARCH.md "LTS cap";
// 12-13 TB/s chip-wide measured here) - the larger of the two paces the tile; DP pays whole waves.
struct GemmPlan { int cg, bn, stream_k, grid; };

// mirrors half_tile_ok<CG, BN, B_MN>() of the kernel
static bool half_tile_ok(int cg, int bn, int b_mn) {
  return b_mn ? ((bn / 2 / cg) % 64 == 0) : ((bn / 2 / cg) % 8 == 0 && (bn / 2) % 16 == 0);
}

static GemmPlan plan_gemm(int M, int N, int K, int b_mn, bool sk_ok, int force_cg, int force_bn) {
  const int ctas = sm_count();
  const int num_kb = ceil_div(K, 64);
  GemmPlan best{1, 128, 0, 1};
  double best_cost = 1e30;
  for (int cg : {2, 1}) {
    if (force_cg && cg != force_cg) continue;
    if (!force_cg && cg == 2 && M <= 128) continue;          // a pair tile would be half empty
    const int workers = ctas / cg;
    for (int bn : {256, 192, 128}) {
      if (b_mn && (bn / cg) % 64 != 0) continue;
      if (force_bn && bn != force_bn) continue;
      const int tiles_m = ceil_div(M, 128 * cg), tiles_n = ceil_div(N, bn);
      const int tiles = tiles_m * tiles_n;
      const double per_kb = fmax(2.0 * bn, (16384.0 + (bn / cg) * 128.0) / 42.0);
      const double per_kb_half = fmax(1.0 * bn, (16384.0 + (bn / cg) * 128.0) / 42.0);
      const double tile_cost = num_kb * per_kb + 1500.0;     // + pipeline fill / non-overlapped epilogue tail
      const double half_cost = num_kb * per_kb_half + 1000.0;
      // data-parallel makespan under the kernel's order: full tiles, then the ragged half-width column, snake rounds
      const bool ragged = half_tile_ok(cg, bn, b_mn) && (N - (tiles_n - 1) * bn) <= bn / 2;
      const int count_full = tiles_m * (tiles_n - (ragged ? 1 : 0));
      const int used = tiles < workers ? tiles : workers;
      double dp_cost = 0.0;
      for (int w = 0; w < used; ++w) {
        double load = 0.0;
        for (int r = 0;; ++r) {
          const int v = r * used + ((r & 1) ? used - 1 - w : w);
          if (v >= tiles) break;
          load += v < count_full ? tile_cost : half_cost;
        }
        dp_cost = fmax(dp_cost, load);
      }
      double cost = dp_cost;
      int sk = 0, grid = used * cg;
      // split mode: whole rounds data-parallel, the tiles of the last partial round cut into k slices (one item per
      // worker).  Worth it when that round would otherwise leave most of the machine idle.
      const int full_rounds = tiles / workers, rem = tiles % workers;
      if (sk_ok && rem > 0) {
        int slices = workers / rem;
        if (slices > num_kb / 2) slices = num_kb / 2;
        if (slices >= 1) {
          const double part = ceil_div(num_kb, slices) * per_kb + 1500.0 + (slices > 1 ? 2500.0 : 0.0);   // + red.add epilogue
          const double sp_cost = full_rounds * tile_cost + part;
          if (sp_cost < 0.95 * dp_cost) {
            cost = sp_cost;
            sk = slices;
            grid = (full_rounds > 0 ? workers : rem * slices) * cg;
          }
        }
      }
      if (cost < best_cost) {
        best_cost = cost;
        best = GemmPlan{cg, bn, sk, grid};
      }
    }
  }
  return best;
}

static int g_force_cg = 0;   // test/debug knob (reed_gemm backend codes 3 / 4): force cta_group::1 / ::2
static int g_force_bn = 0;   // test knob: pin the tile width
void gemm_tcgen05_force_cta_group(int cg) { g_force_cg = cg; }
void gemm_tcgen05_force_bn(int bn) { g_force_bn = bn; }

int gemm_tcgen05(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D, int64_t ldd,
                 int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  // the split mode needs an output that tolerates fp32 atomics: no epilogue, no bias, fp32 D (the weight gradients)
  REED_REQUIRE(ep.bias_grad == nullptr || (ep.kind == kEpiNone && d_dtype == kF32 && ep.bias == nullptr && ep.n_store > 0 &&
                                          ep.n_store < N && ep.n_store % 4 == 0),
               "gemm_tcgen05: a bias-gradient column needs a plain fp32 weight-gradient GEMM");
  const bool sk_ok = ep.kind == kEpiNone && d_dtype == kF32 && ep.bias == nullptr && ep.out2 == nullptr;
  const GemmPlan p = plan_gemm(M, N, K, b_mn, sk_ok, g_force_cg, g_force_bn);
  if (p.stream_k > 1 && !ep.accumulate) {
    // the sliced tiles are the last `rem` of the row-major tile order: zero the output from their first row panel on
    // (whole-round tiles in that range overwrite it with plain stores afterwards - same stream, same kernel order)
    const int workers = sm_count() / p.cg, tiles_n = ceil_div(N, p.bn), tiles = ceil_div(M, 128 * p.cg) * tiles_n;
    const int first = (tiles / workers) * workers;
    const int row0 = (first / tiles_n) * 128 * p.cg;
    if (row0 < M)
      REED_CHECK_CUDA(cudaMemset2DAsync((char*)D + (size_t)row0 * ldd * 4, (size_t)ldd * 4, 0,
                                        (size_t)(ep.bias_grad != nullptr ? ep.n_store : N) * 4, (size_t)(M - row0), st));
  }

  CUtensorMap ma, mb;
  // K-major operand [MN, K] row-major: box = (128 | BN/CG) rows x 64 k.  MN-major operand stored [K, MN]: box = 64 k-rows x 64 mn.
  if (a_mn) { if (make_map(&ma, A, K, M, lda, 64)) return 1; } else { if (make_map(&ma, A, M, K, lda, 128)) return 1; }
  if (b_mn) { if (make_map(&mb, B, K, N, ldb, 64)) return 1; } else { if (make_map(&mb, B, N, K, ldb, p.bn / p.cg)) return 1; }
  // TMA epilogue (epilogue_loop_tma): bf16 D with none / activation / activation-gradient, fp32 D with gate+residual
  static const int tma_epi_on = getenv("REED_TMA_EPI") ? atoi(getenv("REED_TMA_EPI")) : 15;
  EpiMaps em;
  const EpiMaps* emp = nullptr;
  const bool act_kind = ep.kind == kEpiNone || ep.kind == kEpiGelu || ep.kind == kEpiSilu || ep.kind == kEpiDGelu || ep.kind == kEpiDSilu;
  const bool want_o2 = ep.out2 != nullptr && (ep.kind == kEpiGelu || ep.kind == kEpiSilu || ep.kind == kEpiGateRes);
  const bool has_aux = ep.kind == kEpiDGelu || ep.kind == kEpiDSilu || ep.kind == kEpiGateRes;
  // REED_TMA_EPI: bit 0 = activation-gradient kinds (fused operand), bit 1 = activation kinds, bit 2 = plain bf16 stores
  // bit 3 = gate+residual (fp32 D updated in place in a ring of residual boxes, epilogue_loop_tma_gateres) - for short
  // reductions only (attn.proj): the ring takes shared memory from the operand pipeline, and a long main loop (mlp.fc2)
  // loses more to a shallower pipeline than its last tile's epilogue gains (measured equal at K = 4608)
  constexpr int gateres_max_k = 2048;
  const int kind_bit = ep.kind == kEpiGateRes ? 8 : ((ep.kind == kEpiDGelu || ep.kind == kEpiDSilu) ? 1 : ((ep.kind == kEpiGelu || ep.kind == kEpiSilu) ? 2 : 4));
  const bool gate_res = ep.kind == kEpiGateRes && d_dtype == kF32 && ep.rows_per_group % 32 == 0 && K <= gateres_max_k;
  bool tma = (tma_epi_on & kind_bit) && !p.stream_k && !ep.accumulate && ((d_dtype == kBF16 && act_kind) || gate_res) &&
             tma_ok_ptr(D, ldd, d_dtype == kF32 ? 4 : 2) && (!want_o2 || tma_ok_ptr(ep.out2, ep.ld_out2, 2)) &&
             (!has_aux || tma_ok_ptr(ep.aux, ep.ld_aux, ep.kind == kEpiGateRes ? 4 : 2)) &&
             (ep.bias == nullptr || ((uintptr_t)ep.bias & 15) == 0) &&
             (ep.kind != kEpiGateRes || (((uintptr_t)ep.gate & 15) == 0 && ep.ld_gate % 4 == 0));
  if (tma) {
    const int bw = 32;                     // chunk width of the TMA epilogues
    if (make_epi_map(&em.d, D, d_dtype, M, N, ldd, bw)) return 1;
    if (want_o2) { if (make_epi_map(&em.o2, ep.out2, kBF16, M, N, ep.ld_out2, bw)) return 1; } else em.o2 = em.d;
    if (has_aux) { if (make_epi_map(&em.aux, ep.aux, ep.kind == kEpiGateRes ? kF32 : kBF16, M, N, ep.ld_aux, bw)) return 1; } else em.aux = em.d;
    emp = &em;
  }
  if (p.cg == 2) return gemm_tc_launch_cg2(p.bn, a_mn, b_mn, ma, mb, D, ldd, d_dtype, M, N, K, ep, st, p.grid, p.stream_k, emp);
  return gemm_tc_launch_cg1(p.bn, a_mn, b_mn, ma, mb, D, ldd, d_dtype, M, N, K, ep, st, p.grid, p.stream_k, emp);
}

// Grouped operands (see GroupMaps in gemm_tcgen05.cuh).  A: [M, K] K-major bf16, M <= 1024 (cta_group::1 row tiles).
//   mode 0: D[M, groups * n_per_group] = A . [B_0; B_1; ...]^T + bias, B_g [n_per_group, K] K-major (separate allocations)
//   mode 1: D[M, N] (+)= sum_g A_g[M, k_per_group] . B_g[k_per_group, N], B_g MN-major; split over k, fp32 atomics
int gemm_tcgen05_grouped(int mode, const void* const* A, int64_t lda, const void* const* B, int64_t ldb, int groups, int per_group,
                         void* D, int64_t ldd, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  REED_REQUIRE(groups >= 1 && groups <= kMaxGroups, "gemm_grouped: %d groups (1..%d)", groups, kMaxGroups);
  REED_REQUIRE(M >= 1 && M <= 1024, "gemm_grouped: M = %d (the shared input of the grouped linears has at most 1024 rows)", M);
  REED_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldd % 4 == 0, "gemm_grouped: row pitches must keep 16-byte alignment");
  REED_REQUIRE(ep.kind == kEpiNone && ep.out2 == nullptr && ep.bias_grad == nullptr, "gemm_grouped: plain fp32 output only");
  GroupMaps gm;             // ~8 KB, copied into the launch
  gm.mode = mode;
  GemmPlan p;
  if (mode == 0) {
    REED_REQUIRE(N == groups * per_group && per_group % 256 == 0, "gemm_grouped: N = groups x n_per_group, n_per_group %% 256 == 0");
    REED_REQUIRE(!ep.accumulate, "gemm_grouped: the forward form overwrites D");
    gm.per_group = per_group;
    p = plan_gemm(M, N, K, 0, false, 1, 256);
    if (make_map(&gm.a[0], A[0], M, K, lda, 128)) return 1;
    for (int g = 0; g < groups; ++g) {
      REED_REQUIRE(((uintptr_t)B[g] & 15) == 0, "gemm_grouped: operand %d is not 16-byte aligned", g);
      if (make_map(&gm.b[g], B[g], per_group, K, ldb, p.bn)) return 1;
    }
  } else {
    REED_REQUIRE(K == groups * per_group && per_group % 64 == 0, "gemm_grouped: K = groups x k_per_group, k_per_group %% 64 == 0");
    REED_REQUIRE(N % 8 == 0, "gemm_grouped: N %% 8 == 0");
    gm.per_group = per_group / 64;          // in k-blocks
    // split over the reduction (fp32 atomics, order-dependent last bits) only where it pays: short reductions run as plain
    // data-parallel tiles and stay bit-reproducible run to run
    p = plan_gemm(M, N, K, 1, ep.bias == nullptr && K >= 16384, 1, 256);
    if (p.stream_k > 1 && !ep.accumulate) REED_CHECK_CUDA(cudaMemset2DAsync(D, (size_t)ldd * 4, 0, (size_t)N * 4, (size_t)M, st));
    for (int g = 0; g < groups; ++g) {
      REED_REQUIRE((((uintptr_t)A[g] | (uintptr_t)B[g]) & 15) == 0, "gemm_grouped: operand %d is not 16-byte aligned", g);
      if (make_map(&gm.a[g], A[g], M, per_group, lda, 128)) return 1;
      if (make_map(&gm.b[g], B[g], per_group, N, ldb, 64)) return 1;
    }
  }
  return gemm_tc_launch_grouped(p.bn, mode, gm, D, ldd, M, N, K, ep, st, p.grid, p.stream_k);
}

}  // namespace reed
