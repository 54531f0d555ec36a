// Fused attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA),
// forward and backward, for the 256-token patch sequence of the SiT hot path (T = 128 or 256; head_dim 64 or 72).
// Reads Q/K/V straight out of the packed qkv GEMM output [B,T,3,H,hd] through 3-D tensor maps, writes the context
// [B,T,H,hd] and dqkv in the packed layout with TMA stores.  Longer sequences (T = 1024) run the mma.sync kernels of
// attention_mma.cu.
//
// With T <= 256 the whole score row of a query lives in TMEM (128 lanes x 2 x 128 fp32 columns): S = Q K^T is issued
// for both 128-key blocks up front and read from TMEM exactly once by the softmax warps (thread = query row = TMEM
// lane), which write bf16 P into shared memory in the K-major SWIZZLE_128B layout the next MMA (O += P V) consumes.
//
// head_dim 72 is not a multiple of the 64-element swizzle row: every [128 x hd] operand tile is staged as a
// [128 x 64] SWIZZLE_128B tile plus a [128 x 16] SWIZZLE_32B tail whose columns 72..79 are zero-filled by TMA
// (the tensor map's innermost extent is hd, so they are out of bounds).  Contractions over head_dim take 4 + 1
// k-steps; outputs over head_dim are two MMAs (N = 64 and N = 16) into adjacent TMEM columns.
//
//   forward : persistent; work item = 128 query rows of one (batch, head); two items in flight per SM (two slots of
//             operand buffers + 256 TMEM columns, two softmax groups).   TMEM per slot: S0 | S1 (O_j over the dead S_j)
//   dQ      : CTA = 128 query rows; group j owns key block j: S_j = Q K_j^T, dP_j = dO V_j^T,
//             dS_j = P_j (dP_j - delta) -> smem, dQ += dS_j K_j (also emits delta = rowsum(dO * O)).
//   dK/dV   : CTA = 128 key rows, everything transposed (lanes = keys); group i owns query block i:
//             S^T_i = K Q_i^T, dP^T_i = V dO_i^T, dV += P^T_i dO_i, dK += dS^T_i Q_i.
//
// The MMA-issuing warp runs warp-uniform code (descriptors live in uniform registers; one elected lane executes the
// tcgen05 instructions): with 8..64-cycle MMAs the issue path, not the tensor pipe, would otherwise set the pace.
//
// Reference semantics: timm Attention.forward with fused_attn (F.scaled_dot_product_attention, scale hd^-0.5),
// imported at /root/reference/image/models/sit.py:13 and called at sit.py:134; backward = autograd of the same.
#include <cuda.h>
#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace reed {

#ifdef REED_ATTN_TRACE
__device__ unsigned long long g_trace[4][2048];
#define TRACE_DECL(role, cond) int trace_k = 0; const int trace_role = (role); const bool trace_on = blockIdx.x == 0 && (cond)
#define TRACE(tag) do { if (trace_on && trace_k < 2048) g_trace[trace_role][trace_k++] = ((unsigned long long)(tag) << 56) | (clock64() & 0xFFFFFFFFFFFFFFull); } while (0)
extern "C" int reed_debug_trace(void* host) {
  return cudaMemcpyFromSymbol(host, g_trace, sizeof(g_trace)) == cudaSuccess ? 0 : 1;
}
#else
#define TRACE_DECL(role, cond)
#define TRACE(tag)
#endif

namespace {

constexpr int kRows = 128;            // rows of every operand tile (queries or keys per block)

template <int HD> struct Tile {
  static constexpr bool kTail = HD > 64;
  static constexpr int kMain = kRows * 128;                 // [128 x 64] bf16, SWIZZLE_128B
  static constexpr int kTailBytes = kTail ? kRows * 32 : 0; // [128 x 16] bf16, SWIZZLE_32B (cols 72..79 zero)
  static constexpr int kBytes = kMain + kTailBytes;         // 16384 / 20480: multiples of 1024
  static constexpr int kND = kTail ? 80 : 64;               // head_dim as the tensor core sees it
};
constexpr int kPBytes = 2 * kRows * 128;                    // [128 x 128] bf16 as two K-major SWIZZLE_128B tiles

// ---- shared-memory matrix descriptors: constant high word per operand kind, low word = address/16 | LBO/16 << 16 ----
constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) { return (sbo_bytes >> 4) | (1u << 14) | (layout << 29); }
constexpr uint32_t kHiSw128 = desc_hi(1024, 2);   // K-major or MN-major SWIZZLE_128B: 8-row atoms 1024 B apart
constexpr uint32_t kHiSw32 = desc_hi(256, 6);     // SWIZZLE_32B tail: 8-row atoms 256 B apart
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// D[128 x 128] (+)= A[128 x hd] . B[128 x hd]^T, both K-major tiles staged as main + tail.  Warp-uniform; the
// elected lane issues.
template <int HD>
__device__ __forceinline__ void mma_scores(bool leader, uint32_t tmem_d, uint32_t tile_a, uint32_t tile_b) {
  constexpr uint32_t idesc = make_idesc(128, 128, 0, 0);
  const uint32_t la = desc_lo(tile_a), lb = desc_lo(tile_b);
  if (leader) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_bf16<1>(tmem_d, mk_desc(la + ks * 2, kHiSw128), mk_desc(lb + ks * 2, kHiSw128), idesc, ks > 0 ? 1u : 0u);
    if (Tile<HD>::kTail)
      umma_bf16<1>(tmem_d, mk_desc(la + (Tile<HD>::kMain >> 4), kHiSw32), mk_desc(lb + (Tile<HD>::kMain >> 4), kHiSw32),
                   idesc, 1u);
  }
}
// D[128 x hd] (+)= P[128 x 128] . Z[128 x hd]: P = two K-major SWIZZLE_128B tiles written by the softmax warps,
// Z = a TMA-staged tile read MN-major (16 tile rows per k-step); N = 64 main + 16 tail columns.
template <int HD>
__device__ __forceinline__ void mma_accum(bool leader, uint32_t tmem_d, uint32_t p, uint32_t tile_z, bool accumulate) {
  constexpr uint32_t idesc64 = make_idesc(128, 64, 0, 1);
  constexpr uint32_t idesc16 = make_idesc(128, 16, 0, 1);
  const uint32_t lp = desc_lo(p), lz = desc_lo(tile_z);
  if (leader) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t da = mk_desc(lp + (ks >> 2) * (kRows * 128 >> 4) + (ks & 3) * 2, kHiSw128);
      umma_bf16<1>(tmem_d, da, mk_desc(lz + ks * (2048 >> 4), kHiSw128), idesc64, (accumulate || ks > 0) ? 1u : 0u);
      if (Tile<HD>::kTail)
        umma_bf16<1>(tmem_d + 64, da, mk_desc(lz + (Tile<HD>::kMain >> 4) + ks * (512 >> 4), kHiSw32), idesc16,
                     (accumulate || ks > 0) ? 1u : 0u);
    }
  }
}
__device__ __forceinline__ void commit_if(bool leader, uint64_t* bar) {
  if (leader) umma_commit<1>(bar);
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// one [128 x hd] tile: rows row0.., "head" column block hcol of the packed tensor
template <int HD>
__device__ __forceinline__ void load_tile(const CUtensorMap* main, const CUtensorMap* tail, uint64_t* bar, uint32_t dst,
                                          int hcol, int row0) {
  tma_load_3d(main, bar, dst, 0, hcol, row0);
  if (Tile<HD>::kTail) tma_load_3d(tail, bar, dst + Tile<HD>::kMain, 64, hcol, row0);
}
// store a staged [128 x hd] tile (main SWIZZLE_128B at `src`, 8-column tail rows of 16 B at src + kMain)
template <int HD>
__device__ __forceinline__ void store_tile(const CUtensorMap* main, const CUtensorMap* tail8, uint32_t src, int hcol, int row0) {
  tma_store_3d(main, src, 0, hcol, row0);
  if (Tile<HD>::kTail) tma_store_3d(tail8, src + Tile<HD>::kMain, 64, hcol, row0);
}

// 32 / 16 consecutive TMEM columns of this thread's lane, no wait (pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
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
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void sts128(uint32_t addr, const float* v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2(v[0], v[1])), "r"(pack2(v[2], v[3])),
               "r"(pack2(v[4], v[5])), "r"(pack2(v[6], v[7]))
               : "memory");
}
// 32 values of row `row` (columns c0..c0+31 of a [128 x 128] K-major SWIZZLE_128B pair of tiles) -> shared memory.
// 16-byte chunk c of a row sits at chunk position c ^ (row & 7): the 8 lanes of a quarter-warp hit 8 distinct
// positions, so the stores are bank-conflict free.
__device__ __forceinline__ void store_p32(uint32_t base, int row, int c0, const float* v) {
  const uint32_t tile = base + (c0 >> 6) * (kRows * 128) + row * 128;
  const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) sts128(tile + (((chunk0 + q) ^ (row & 7)) << 4), v + 8 * q);
}
// 8 output values (columns 8*chunk..) of row `row` into a TMA-store staging tile: chunks 0..7 go to the
// SWIZZLE_128B main tile, chunk 8 to the 16-byte-per-row tail
template <int HD>
__device__ __forceinline__ void stage_out8(uint32_t tile, int row, int chunk, const float* v) {
  if (chunk < 8) sts128(tile + row * 128 + ((chunk ^ (row & 7)) << 4), v);
  else sts128(tile + Tile<HD>::kMain + row * 16, v);
}
// read 8 bf16 (16-byte chunk `chunk`, 0..7 main, 8..9 tail) of row `row` of a TMA-staged [128 x hd] tile
template <int HD>
__device__ __forceinline__ uint4 load_tile_chunk(uint32_t tile, int row, int chunk) {
  uint32_t addr;
  if (chunk < 8) addr = tile + row * 128 + ((chunk ^ (row & 7)) << 4);
  else addr = tile + Tile<HD>::kMain + row * 32 + ((((chunk - 8) ^ ((row >> 2) & 1))) << 4);
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc = fmaf(__low2float(pa[i]), __low2float(pb[i]), acc);
    acc = fmaf(__high2float(pa[i]), __high2float(pb[i]), acc);
  }
  return acc;
}
// columns [c_begin, c_end) (multiples of 8, c_end <= 72) of this thread's accumulator row -> scaled bf16 staging tile
template <int HD>
__device__ __forceinline__ void stage_acc_row(uint32_t taddr, float mul, uint32_t tile, int row, int c_begin, int c_end) {
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    float v[16];
    tmem_ld16_nowait(taddr + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= mul;
    stage_out8<HD>(tile, row, c0 >> 3, v);
    if (c0 + 8 < c_end) stage_out8<HD>(tile, row, (c0 >> 3) + 1, v + 8);
  }
}

struct AttnMaps {
  CUtensorMap qkv_main, qkv_tail;   // [B*T, 3H, hd] loads
  CUtensorMap o_main, o_tail;       // [B*T, H, hd]   saved context (loaded by the dQ kernel)
  CUtensorMap do_main, do_tail;     // [B*T, H, hd]
  CUtensorMap out_main, out_tail8;  // stores: forward -> o, backward -> dqkv (tail8 = 8-column box, no swizzle)
};

// ------------------------------------------------------------------------------------------------------
// forward - persistent, two work items in flight per SM
//
// A work item is one 128-query tile of one (batch, head).  Every CTA owns two slots (shared-memory operand buffers +
// 256 TMEM columns each); its n-th item lives in slot n & 1 and is handled by softmax group n & 1 (warps 0-3 /
// 4-7).  Warp 8 streams the operands of the next item into a slot as soon as its previous item has left it; warp 9
// polls the barriers of both slots and issues whichever MMA batch is ready (S = Q K^T, O_0 = P_0 V_0, O_1 = P_1 V_1),
// so one group exponentiates while the other waits for its MMAs.  Scores are read from TMEM once: block 0 is
// exponentiated against its own row maximum into O_0, block 1 against the final maximum into O_1, and the epilogue
// combines O = (O_0 2^(m0-m) + O_1) / l, stages the bf16 rows in shared memory and TMA-stores them.
// ------------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 320;

template <int HD>
struct FwdSmem {
  static constexpr int kSlot = 5 * Tile<HD>::kBytes;          // Q | K0 K1 (P, then the output staging, alias) | V0 V1
  static constexpr int kBars = 2 * kSlot;                     // byte offset of the barrier block
  static constexpr int kTotal = 1024 + 2 * kSlot + 256;
};

template <int HD>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_tc5_fwd_kernel(const __grid_constant__ AttnMaps maps, float* __restrict__ lse, int T, int H, int num_items,
                    float scale_log2) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  using TL = Tile<HD>;
  using SM = FwdSmem<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBars);
  // per slot s (stride 13): [0] Q landed, [1..2] K_j, [3..4] V_j, [5..6] S_j complete, [7..8] P_j written,
  // [9..10] O_j = P_j V_j retired, [11] accumulators read out (TMEM slot free), [12] output store has read the slot
  auto BAR = [&](int slot, int k) { return bars + slot * 13 + k; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = T / kRows;
  const int nq = T / kRows;
  const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 256) {
    for (int s = 0; s < 2; ++s) {
      for (int k = 0; k < 7; ++k) mbar_init(BAR(s, k), 1);
      mbar_init(BAR(s, 7), 128);
      mbar_init(BAR(s, 8), 128);
      mbar_init(BAR(s, 9), 1);
      mbar_init(BAR(s, 10), 1);
      mbar_init(BAR(s, 11), 128);
      mbar_init(BAR(s, 12), 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.qkv_main);
    tma_prefetch_desc(&maps.out_main);
    if (TL::kTail) {
      tma_prefetch_desc(&maps.qkv_tail);
      tma_prefetch_desc(&maps.out_tail8);
    }
  }
  if (warp == 9) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      // ---- TMA producer ----
      TRACE_DECL(3, true);
      for (int n = 0; n < my_items; ++n) {
        const int s = n & 1;
        const uint32_t ph = (uint32_t)(n >> 1) & 1u;
        TRACE(1);
        if (n >= 2) mbar_wait(BAR(s, 12), ph ^ 1u);     // item n-2 has left the slot (its output store has read it)
        TRACE(2);
        const int it = (int)blockIdx.x + n * (int)gridDim.x;
        const int qb = it % nq, h = (it / nq) % H, b = it / (nq * H);
        const uint32_t sQ = sbase + s * SM::kSlot, sK = sQ + TL::kBytes, sV = sK + 2 * TL::kBytes;
        mbar_expect_tx(BAR(s, 0), TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, BAR(s, 0), sQ, h, b * T + qb * kRows);
        for (int j = 0; j < nblk; ++j) {
          mbar_expect_tx(BAR(s, 1 + j), TL::kBytes);
          load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, BAR(s, 1 + j), sK + j * TL::kBytes, H + h, b * T + j * kRows);
        }
        for (int j = 0; j < nblk; ++j) {
          mbar_expect_tx(BAR(s, 3 + j), TL::kBytes);
          load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, BAR(s, 3 + j), sV + j * TL::kBytes, 2 * H + h, b * T + j * kRows);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---- MMA issuer: warp-uniform polling state machine over the two slots ----
    const bool leader = elect_one();
    TRACE_DECL(2, lane == 0);
    int cur[2] = {0, 1};          // next item of each slot
    int stage[2] = {0, 0};        // 0: S pending, 1: O_0 pending, 2: O_1 pending
    int remaining = my_items;
    while (remaining > 0) {
      bool progress = false;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int n = cur[s];
        if (n >= my_items) continue;
        const uint32_t ph = (uint32_t)(n >> 1) & 1u;
        const uint32_t sQ = sbase + s * SM::kSlot, sK = sQ + TL::kBytes, sV = sK + 2 * TL::kBytes;
        const uint32_t d = tmem + s * 256;
        if (stage[s] == 0) {
          bool ready = (n < 2 || mbar_test(BAR(s, 11), ph ^ 1u)) && mbar_test(BAR(s, 0), ph) && mbar_test(BAR(s, 1), ph);
          if (ready && nblk > 1) ready = mbar_test(BAR(s, 2), ph);
          if (!__all_sync(0xffffffffu, ready)) continue;
          tc_fence_after();
          for (int j = 0; j < nblk; ++j) {
            mma_scores<HD>(leader, d + j * 128, sQ, sK + j * TL::kBytes);
            commit_if(leader, BAR(s, 5 + j));
          }
          TRACE(3);
          stage[s] = 1;
          progress = true;
        } else {
          const int j = stage[s] - 1;
          if (!__all_sync(0xffffffffu, mbar_test(BAR(s, 3 + j), ph) && mbar_test(BAR(s, 7 + j), ph))) continue;
          tc_fence_after();
          mma_accum<HD>(leader, d + j * 128, sK, sV + j * TL::kBytes, false);     // P aliases the K tiles
          commit_if(leader, BAR(s, 9 + j));
          TRACE(6);
          if (j + 1 < nblk) stage[s] = 2;
          else { stage[s] = 0; cur[s] = n + 2; --remaining; }
          progress = true;
        }
      }
      if (!progress) __nanosleep(32);
    }
    __syncwarp();
  } else {
    // ---- softmax groups: thread = query row = TMEM lane ----
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + g * 256;
    const uint32_t sP = sbase + g * SM::kSlot + TL::kBytes;       // P, then the output staging tile
    TRACE_DECL(g, (warp & 3) == 0 && lane == 0);
    for (int n = g; n < my_items; n += 2) {
      const uint32_t ph = (uint32_t)(n >> 1) & 1u;
      const int it = (int)blockIdx.x + n * (int)gridDim.x;
      const int qb = it % nq, h = (it / nq) % H, b = it / (nq * H);
      float m0c, mc, l0, l1 = 0.f;
      TRACE(1);
      for (int j = 0; j < nblk; ++j) mbar_wait(BAR(g, 5 + j), ph);   // P_0 overwrites K_0 and K_1: both S done
      tc_fence_after();
      TRACE(2);
      {
        float v[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32_nowait(trow + c * 32, v + c * 32);
        tmem_wait_ld();
        float m = v[0];
#pragma unroll
        for (int i = 1; i < 128; ++i) m = fmaxf(m, v[i]);
        m0c = m * scale_log2;
        float la = 0.f, lb = 0.f;
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          v[i] = ex2(fmaf(v[i], scale_log2, -m0c));
          v[i + 1] = ex2(fmaf(v[i + 1], scale_log2, -m0c));
          la += v[i];
          lb += v[i + 1];
        }
        l0 = la + lb;
#pragma unroll
        for (int c = 0; c < 4; ++c) store_p32(sP, row, c * 32, v + c * 32);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(BAR(g, 7));
      TRACE(3);
      mc = m0c;
      if (nblk > 1) {
        float v[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32_nowait(trow + 128 + c * 32, v + c * 32);
        tmem_wait_ld();
        float m = v[0];
#pragma unroll
        for (int i = 1; i < 128; ++i) m = fmaxf(m, v[i]);
        mc = fmaxf(m0c, m * scale_log2);
        float la = 0.f, lb = 0.f;
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          v[i] = ex2(fmaf(v[i], scale_log2, -mc));
          v[i + 1] = ex2(fmaf(v[i + 1], scale_log2, -mc));
          la += v[i];
          lb += v[i + 1];
        }
        l1 = la + lb;
        TRACE(4);
        mbar_wait(BAR(g, 9), ph);                 // the MMA has finished reading P_0
        TRACE(5);
#pragma unroll
        for (int c = 0; c < 4; ++c) store_p32(sP, row, c * 32, v + c * 32);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(BAR(g, 8));
        TRACE(6);
      }
      mbar_wait(BAR(g, 9 + nblk - 1), ph);
      tc_fence_after();
      TRACE(7);
      const float a0 = ex2(m0c - mc);
      const float l = fmaf(l0, a0, l1);
      const float inv = 1.f / l, w0 = a0 * inv;
      const int tok = qb * kRows + row;
      lse[((int64_t)b * H + h) * T + tok] = (mc + log2f(l)) * 0.6931471805599453f;
#pragma unroll
      for (int c0 = 0; c0 < TL::kND; c0 += 16) {
        float x0[16], x1[16];
        tmem_ld16_nowait(trow + c0, x0);
        if (nblk > 1) tmem_ld16_nowait(trow + 128 + c0, x1);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) x0[i] = nblk > 1 ? fmaf(x0[i], w0, x1[i] * inv) : x0[i] * w0;
        stage_out8<HD>(sP, row, c0 >> 3, x0);
        if (c0 + 8 < HD) stage_out8<HD>(sP, row, (c0 >> 3) + 1, x0 + 8);
      }
      tc_fence_before();
      mbar_arrive(BAR(g, 11));                    // TMEM slot free
      fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if ((warp & 3) == 0 && elect_one()) {
        store_tile<HD>(&maps.out_main, &maps.out_tail8, sP, h, b * T + qb * kRows);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(BAR(g, 12));                  // the slot's shared memory may be refilled
      }
      TRACE(8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------
// backward, part 1: dQ and delta.  CTA = 128 query rows; group j (warps 4j..4j+3) owns key block j.
// ------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 288;      // warps 0-7: two elementwise groups; warp 8: TMA + MMA

template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_tc5_dq_kernel(const __grid_constant__ AttnMaps maps, const float* __restrict__ lse, float* __restrict__ delta,
                   int T, int H, float scale, float scale_log2) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  using TL = Tile<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sQ = smem_u32(smem);             // Q; the dQ staging tile once the score MMAs have retired
  const uint32_t sDO = sQ + TL::kBytes;
  const uint32_t sK = sDO + TL::kBytes;           // 2 tiles
  const uint32_t sV = sK + 2 * TL::kBytes;        // 2 tiles
  const uint32_t sDS = sV + 2 * TL::kBytes;       // 2 x kPBytes; the saved context O is staged in the second one first
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TL::kBytes + 2 * kPBytes);
  uint64_t* bar_q = bars;          // Q, dO, O landed
  uint64_t* bar_kv = bars + 1;     // [2] K_j, V_j landed
  uint64_t* bar_s = bars + 3;      // [2] S_j and dP_j complete
  uint64_t* bar_ds = bars + 5;     // [2] dS_j written
  uint64_t* bar_dq = bars + 7;     // dQ complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nblk = T / kRows;
  const int row0 = b * T + qb * kRows;

  if (threadIdx.x == 256) {
    mbar_init(bar_q, 1);
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bar_kv[j], 1);
      mbar_init(&bar_s[j], 1);
      mbar_init(&bar_ds[j], 128);
    }
    mbar_init(bar_dq, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(bar_q, 3 * TL::kBytes);
      load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bar_q, sQ, h, row0);
      load_tile<HD>(&maps.do_main, &maps.do_tail, bar_q, sDO, h, row0);
      load_tile<HD>(&maps.o_main, &maps.o_tail, bar_q, sDS + kPBytes, h, row0);
      for (int j = 0; j < nblk; ++j) {
        mbar_expect_tx(&bar_kv[j], 2 * TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_kv[j], sK + j * TL::kBytes, H + h, b * T + j * kRows);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_kv[j], sV + j * TL::kBytes, 2 * H + h, b * T + j * kRows);
      }
    }
    __syncwarp();
    mbar_wait(bar_q, 0);
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&bar_kv[j], 0);
      tc_fence_after();
      mma_scores<HD>(leader, tmem + j * 256, sQ, sK + j * TL::kBytes);            // S_j = Q K_j^T
      mma_scores<HD>(leader, tmem + j * 256 + 128, sDO, sV + j * TL::kBytes);     // dP_j = dO V_j^T
      commit_if(leader, &bar_s[j]);
    }
    for (int j = 0; j < nblk; ++j) {                  // dQ += dS_j K_j  (columns 0..kND-1, over the dead S_0)
      mbar_wait(&bar_ds[j], 0);
      tc_fence_after();
      mma_accum<HD>(leader, tmem, sDS + j * kPBytes, sK + j * TL::kBytes, j > 0);
    }
    commit_if(leader, bar_dq);
    __syncwarp();
  } else {
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int tok = qb * kRows + row;
    const int64_t sidx = ((int64_t)b * H + h) * T + tok;
    const float ls = lse[sidx] * 1.4426950408889634f;     // exp2 domain
    // delta = rowsum(dO * O) from the staged tiles (both groups need it; group 0 publishes it)
    mbar_wait(bar_q, 0);
    float dl = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) dl += dot8(load_tile_chunk<HD>(sDO, row, c), load_tile_chunk<HD>(sDS + kPBytes, row, c));
    if (g == 0) delta[sidx] = dl;
    // every thread is done with the staged O before dS_1 overwrites it
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (g < nblk) {
      mbar_wait(&bar_s[g], 0);
      tc_fence_after();
      const uint32_t ds = sDS + g * kPBytes;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 64) {
        float s[64], dp[64];
        tmem_ld32_nowait(trow + g * 256 + c0, s);
        tmem_ld32_nowait(trow + g * 256 + c0 + 32, s + 32);
        tmem_ld32_nowait(trow + g * 256 + 128 + c0, dp);
        tmem_ld32_nowait(trow + g * 256 + 128 + c0 + 32, dp + 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float p = ex2(fmaf(s[i], scale_log2, -ls));
          s[i] = p * (dp[i] - dl);                 // dS without the softmax scale (applied to dQ at the end)
        }
        store_p32(ds, row, c0, s);
        store_p32(ds, row, c0 + 32, s + 32);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar_ds[g]);
    }
    mbar_wait(bar_dq, 0);
    tc_fence_after();
    // stage dQ (group 0: columns 0..31 and the tail, group 1: columns 32..63) over the dead Q tile and TMA-store it
    stage_acc_row<HD>(trow, scale, sQ, row, g * 32, g * 32 + 32);
    if (TL::kTail && g == 0) stage_acc_row<HD>(trow, scale, sQ, row, 64, 72);
    fence_proxy_async();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 0 && elect_one()) {
      store_tile<HD>(&maps.out_main, &maps.out_tail8, sQ, h, row0);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<1>(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------
// backward, part 2: dK and dV.  CTA = 128 key rows; lanes = keys, columns = queries; group i owns query block i.
// ------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_tc5_dkv_kernel(const __grid_constant__ AttnMaps maps, const float* __restrict__ lse, const float* __restrict__ delta,
                    int T, int H, float scale, float scale_log2) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  using TL = Tile<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sK = smem_u32(smem);             // K | V: dead once the four score MMAs retire -> dS^T_1 lives here
  const uint32_t sV = sK + TL::kBytes;
  const uint32_t sQ = sV + TL::kBytes;            // 2 tiles; dV / dK staging once the accumulate MMAs have retired
  const uint32_t sDO = sQ + 2 * TL::kBytes;       // 2 tiles
  const uint32_t sBuf = sDO + 2 * TL::kBytes;     // 3 x kPBytes: P^T_0, dS^T_0, P^T_1
  float* sL = reinterpret_cast<float*>(smem + 6 * TL::kBytes + 3 * kPBytes);   // [256] lse (exp2 domain)
  float* sD = sL + 256;                                                          // [256] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
  uint64_t* bar_kv = bars;         // K, V landed
  uint64_t* bar_q = bars + 1;      // [2] Q_i, dO_i landed
  uint64_t* bar_s = bars + 3;      // [2] S^T_i, dP^T_i complete
  uint64_t* bar_p = bars + 5;      // [2] P^T_i, dS^T_i written
  uint64_t* bar_acc = bars + 7;    // dV, dK complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nblk = T / kRows;
  const int row0 = b * T + kb * kRows;

  if (threadIdx.x == 256) {
    mbar_init(bar_kv, 1);
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bar_q[j], 1);
      mbar_init(&bar_s[j], 1);
      mbar_init(&bar_p[j], 128);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<1>(tmem_slot, 512);
  // per-query statistics of this (batch, head)
  for (int i = threadIdx.x; i < T; i += kBwdThreads) {
    sL[i] = lse[((int64_t)b * H + h) * T + i] * 1.4426950408889634f;
    sD[i] = delta[((int64_t)b * H + h) * T + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    const bool leader = elect_one();
    TRACE_DECL(2, lane == 0);
    TRACE(1);
    if (leader) {
      mbar_expect_tx(bar_kv, 2 * TL::kBytes);
      load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bar_kv, sK, H + h, row0);
      load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bar_kv, sV, 2 * H + h, row0);
      for (int i = 0; i < nblk; ++i) {
        mbar_expect_tx(&bar_q[i], 2 * TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_q[i], sQ + i * TL::kBytes, h, b * T + i * kRows);
        load_tile<HD>(&maps.do_main, &maps.do_tail, &bar_q[i], sDO + i * TL::kBytes, h, b * T + i * kRows);
      }
    }
    __syncwarp();
    mbar_wait(bar_kv, 0);
    TRACE(2);
    for (int i = 0; i < nblk; ++i) {
      mbar_wait(&bar_q[i], 0);
      TRACE(3);
      tc_fence_after();
      mma_scores<HD>(leader, tmem + i * 256, sK, sQ + i * TL::kBytes);            // S^T_i = K Q_i^T
      mma_scores<HD>(leader, tmem + i * 256 + 128, sV, sDO + i * TL::kBytes);     // dP^T_i = V dO_i^T
      commit_if(leader, &bar_s[i]);
    }
    for (int i = 0; i < nblk; ++i) {                  // dV -> columns 0.. (dead S^T_0), dK -> columns 128.. (dead dP^T_0)
      mbar_wait(&bar_p[i], 0);
      TRACE(4);
      tc_fence_after();
      const uint32_t pt = sBuf + (i == 0 ? 0 : 2 * kPBytes), dst = i == 0 ? sBuf + kPBytes : sK;
      mma_accum<HD>(leader, tmem, pt, sDO + i * TL::kBytes, i > 0);
      mma_accum<HD>(leader, tmem + 128, dst, sQ + i * TL::kBytes, i > 0);
      TRACE(5);
    }
    commit_if(leader, bar_acc);
    __syncwarp();
  } else {
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    TRACE_DECL(g, (warp & 3) == 0 && lane == 0);
    TRACE(1);
    if (g < nblk) {
      // dS^T_1 overwrites K and V: every score MMA (both blocks) must have retired
      mbar_wait(&bar_s[g], 0);
      if (g == 1) mbar_wait(&bar_s[0], 0);
      tc_fence_after();
      TRACE(2);
      const uint32_t pt = sBuf + (g == 0 ? 0 : 2 * kPBytes), dst = g == 0 ? sBuf + kPBytes : sK;
      const float* L = sL + g * kRows;
      const float* Dl = sD + g * kRows;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float s[32], dp[32];
        tmem_ld32_nowait(trow + g * 256 + c0, s);
        tmem_ld32_nowait(trow + g * 256 + 128 + c0, dp);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(L + c0 + q);      // same address in every lane: broadcast
          const float4 d4 = *reinterpret_cast<const float4*>(Dl + c0 + q);
          const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p = ex2(fmaf(s[q + e], scale_log2, -lq[e]));
            s[q + e] = p;
            dp[q + e] = p * (dp[q + e] - dq[e]);
          }
        }
        store_p32(pt, row, c0, s);
        store_p32(dst, row, c0, dp);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar_p[g]);
      TRACE(3);
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    TRACE(4);
    // group 0 stages and stores dV (over Q_0), group 1 dK (over Q_1)
    const uint32_t stage = sQ + g * TL::kBytes;
    stage_acc_row<HD>(trow + g * 128, g == 0 ? 1.f : scale, stage, row, 0, HD);
    fence_proxy_async();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    if ((warp & 3) == 0 && elect_one()) {
      store_tile<HD>(&maps.out_main, &maps.out_tail8, stage, (g == 0 ? 2 * H : H) + h, row0);
      tma_store_commit();
      tma_store_wait_read();
    }
    TRACE(5);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<1>(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// [rows, heads, hd] bf16, hd contiguous: box = 128 rows x 1 head x cols.  kind 0: 64 columns, 128B swizzle (main);
// 1: 16 columns, 32B swizzle (operand tail, columns hd.. zero-filled); 2: 8 columns, no swizzle (output tail)
int make_map3(CUtensorMap* map, const void* ptr, int64_t rows, int heads, int hd, int kind) {
  EncodeTiledFn enc = encode_fn();
  REED_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)hd, (cuuint64_t)heads, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)hd * 2, (cuuint64_t)heads * hd * 2};
  cuuint32_t box[3] = {kind == 0 ? 64u : (kind == 1 ? 16u : 8u), 1u, (cuuint32_t)kRows};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUtensorMapSwizzle sw = kind == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : (kind == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REED_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d) rows=%lld heads=%d hd=%d kind=%d", (int)r,
               (long long)rows, heads, hd, kind);
  return 0;
}

template <int HD> constexpr int fwd_smem() { return FwdSmem<HD>::kTotal; }
template <int HD> constexpr int dq_smem() { return 1024 + 6 * Tile<HD>::kBytes + 2 * kPBytes + 128; }
template <int HD> constexpr int dkv_smem() { return 1024 + 6 * Tile<HD>::kBytes + 3 * kPBytes + 2048 + 128; }

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = kNumSMs;
  }
  return sms;
}

template <int HD>
int fwd_launch(const void* qkv, void* o, float* lse, int B, int T, int H, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_tc5_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem<HD>()));
    done = true;
  }
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int64_t rows = (int64_t)B * T;
  if (make_map3(&maps.qkv_main, qkv, rows, 3 * H, HD, 0)) return 1;
  if (make_map3(&maps.out_main, o, rows, H, HD, 0)) return 1;
  if (Tile<HD>::kTail) {
    if (make_map3(&maps.qkv_tail, qkv, rows, 3 * H, HD, 1)) return 1;
    if (make_map3(&maps.out_tail8, o, rows, H, HD, 2)) return 1;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  const int items = (T / kRows) * H * B;
  const int grid = items < sm_count() ? items : sm_count();
  attn_tc5_fwd_kernel<HD><<<grid, kFwdThreads, fwd_smem<HD>(), st>>>(maps, lse, T, H, items, scale_log2);
  REED_LAUNCH_CHECK();
  return 0;
}

template <int HD>
int bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta, int B, int T,
               int H, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_tc5_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq_smem<HD>()));
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_tc5_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv_smem<HD>()));
    done = true;
  }
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int64_t rows = (int64_t)B * T;
  if (make_map3(&maps.qkv_main, qkv, rows, 3 * H, HD, 0)) return 1;
  if (make_map3(&maps.o_main, o, rows, H, HD, 0)) return 1;
  if (make_map3(&maps.do_main, d_o, rows, H, HD, 0)) return 1;
  if (make_map3(&maps.out_main, dqkv, rows, 3 * H, HD, 0)) return 1;
  if (Tile<HD>::kTail) {
    if (make_map3(&maps.qkv_tail, qkv, rows, 3 * H, HD, 1)) return 1;
    if (make_map3(&maps.o_tail, o, rows, H, HD, 1)) return 1;
    if (make_map3(&maps.do_tail, d_o, rows, H, HD, 1)) return 1;
    if (make_map3(&maps.out_tail8, dqkv, rows, 3 * H, HD, 2)) return 1;
  }
  const float scale = 1.f / sqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(T / kRows, H, B);
  attn_tc5_dq_kernel<HD><<<grid, kBwdThreads, dq_smem<HD>(), st>>>(maps, lse, delta, T, H, scale, scale_log2);
  attn_tc5_dkv_kernel<HD><<<grid, kBwdThreads, dkv_smem<HD>(), st>>>(maps, lse, delta, T, H, scale, scale_log2);
  REED_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool attn_tc5_supported(int T, int hd) { return (T == 128 || T == 256) && (hd == 64 || hd == 72); }

int attn_tc5_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int hd, cudaStream_t st) {
  if (B == 0) return 0;
  if (hd == 64) return fwd_launch<64>(qkv, o, lse, B, T, H, st);
  if (hd == 72) return fwd_launch<72>(qkv, o, lse, B, T, H, st);
  return fail("tcgen05 attention: head_dim %d unsupported", hd);
}

int attn_tc5_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta, int B,
                 int T, int H, int hd, cudaStream_t st) {
  if (B == 0) return 0;
  if (hd == 64) return bwd_launch<64>(qkv, o, d_o, lse, dqkv, delta, B, T, H, st);
  if (hd == 72) return bwd_launch<72>(qkv, o, d_o, lse, dqkv, delta, B, T, H, st);
  return fail("tcgen05 attention: head_dim %d unsupported", hd);
}

}  // namespace reed
