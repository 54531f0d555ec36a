// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands staged in
// shared memory by TMA with the 128-byte swizzle, persistent over output tiles, warp-specialised:
//   warp 0    : TMA producer          (one elected lane)
//   warp 1    : tcgen05.mma issuer    (one elected lane), commits free smem stages / publish accumulators
//   warp 2    : TMEM allocator
//   warps 3-10: epilogue - tcgen05.ld the 128 x BN accumulator, transpose through smem so global accesses are
//               row-coalesced, apply the fused epilogue (bias / GELU / SiLU / gate*y+residual / act') and store.
// Two accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
//
// CG = 2 (`cta_group::2`): a CTA pair on one TPC computes a 256 x BN tile.  Each CTA stages its own 128 rows of A
// and its own BN/2 rows of B (half the L2->SM operand traffic per FLOP of the single-CTA kernel, which is what
// bounds these GEMMs on B200), the leader CTA issues the M = 256 MMAs for both, every CTA drains the 128
// accumulator rows that live in its own TMEM.
//
//   D[M,N] = epi( A[M,K] . B[N,K]^T )        A, B bf16; each either K-major or MN-major in global memory:
//   forward  y  = x W^T      : A = x  (K-major),  B = W  (K-major)
//   dgrad    dx = dy W       : A = dy (K-major),  B = W  (MN-major: stored [N_contract, K_out])
//   wgrad    dW = dy^T x     : A = dy (MN-major), B = x  (MN-major)
// This covers the qkv / proj / fc1 / fc2 / adaLN / projector linears of /root/reference/image/models/sit.py
// (timm Attention.qkv/proj, Mlp.fc1/fc2 at sit.py:114-124; adaLN 125-128; build_mlp 17-24) and their backward.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace reed {

constexpr int BM = 128;         // accumulator rows per CTA (UMMA M = 128 x CG)
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 352;   // warps 0-2: TMA / MMA / TMEM alloc; warps 3-10: epilogue (184 registers per thread)
constexpr int kFirstEpiWarp = 3;
constexpr int kEpiWarps = 8;
constexpr int kStageCols = 32;  // accumulator columns moved per tcgen05.ld
constexpr int kStagePitch = 32; // floats; 16-byte chunk j of row r sits at chunk j ^ (r & 7): float4 accesses stay conflict-free
                                // without padding (4 KB per warp - the 32 KB region leaves the 256-wide tiles a sixth stage)

constexpr int kEpiAccum = 6;    // internal: kEpiNone with ep.accumulate (D += acc), fp32 D
constexpr int kEpiAtomic = 7;   // internal: stream-K partial tile, red.add into fp32 D

template <int CG, int BN>
struct GemmCfg {
  static constexpr int kBNL = BN / CG;            // B rows (output columns) staged by each CTA
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBNL * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 32 * kStagePitch * 4;   // epilogue region of the register/LSU epilogue
  static constexpr int kTmaEpiBytes = kEpiWarps * 4096;                    // epilogue region of the TMA epilogue ...
  static constexpr int kTmaEpiAuxBytes = kEpiWarps * 8192;                 // ... with a fused global operand
  static constexpr int kBarBytes = 1024;
  // shared memory = 1024 (alignment slack) + stages * kStageBytes + epilogue region + barriers
  static constexpr int stages_for(int epi_bytes) {
    int s = (227 * 1024 - 1024 - epi_bytes - kBarBytes) / kStageBytes;
    return s > 8 ? 8 : s;
  }
  static constexpr int smem_bytes(int stages, int epi_bytes) { return 1024 + stages * kStageBytes + epi_bytes + kBarBytes; }
  static constexpr int kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
};

// Tensor maps of the TMA epilogue: D, out2 (bf16) and the fused global operand, all as [32 rows x 32 columns] boxes
// (one epilogue warp's share of a chunk) with the 64-byte (bf16) / 128-byte (fp32) swizzle.
struct EpiMaps { CUtensorMap d, o2, aux; };

// Grouped operands (the adaLN-Zero modulation linears of ALL blocks as one launch, sit.py:125-133): up to kMaxGroups
// weight matrices that live in separate allocations, presented to the kernel as one virtual operand.
//   mode 0 (forward, mod_g = c W_g^T): the virtual B is the groups' [n_per_group, K] matrices stacked along N; a column
//           tile lies in one group (n_per_group % BN == 0) and the producer picks that group's tensor map.
//   mode 1 (dgrad, dc = sum_g dmod_g W_g): the groups are concatenated along the reduction; k-block kb belongs to group
//           kb / kb_per_group and both operands come from that group's maps.
// The maps travel as a kernel parameter (__grid_constant__, ~8 KB), so the call is capturable in a CUDA graph.
constexpr int kMaxGroups = 32;
struct GroupMaps { CUtensorMap a[kMaxGroups], b[kMaxGroups]; int mode, per_group; };
struct NoGroups { int unused; };

// Work distribution.  Data-parallel: output tiles round-robin over the persistent CTAs (pairs), rounds in lock step,
// so the CTAs running at any moment read the same k range of neighbouring tiles and L2 serves each operand line to
// several SMs at once.  Split mode (fp32 outputs that tolerate atomics, i.e. the weight-gradient GEMMs, whose tile
// count does not fill the machine evenly): the whole rounds run data-parallel with plain stores; the tiles of the
// last, partial round are cut into `slices` equal k ranges, one (tile, slice) item per worker, neighbouring workers
// on neighbouring tiles of the SAME slice, and added into the zeroed / accumulating output with red.global.add.
// (A contiguous stream-K split gave every worker its own k phase: no two SMs wanted the same line at the same
// time and the operand fetch fell back to the L2 data-array rate.)
struct Seg { int tile, kb0, kb1, half, atomic; };   // half: the tile is the ragged last column tile, computed BN/2 wide
struct Sched {
  int slices, num_tiles, num_kb, tiles_n, ragged, round;
  int worker, stride, colmajor, rem_done;
  // `worker` = index of this CTA (CG = 1) or CTA pair (CG = 2) among `workers`.
  // Data-parallel order: all full-width tiles first (row-major), then the ragged last-column tiles (when N leaves a
  // remainder of at most BN/2 they are computed with a BN/2-wide MMA and cost half a tile); rounds alternate
  // direction over the workers (snake), so the half tiles of the last rounds land on the workers that got one
  // tile less - e.g. N = 1152, BN = 256, M = 8192: 128 full + 32 half tiles on 74 pairs take 2 tile times, not 3.
  __device__ Sched(int slices_, int tiles_m, int tiles_n_, int ragged_, int num_kb_, int worker_, int workers, int colmajor_ = 0)
      : slices(slices_), num_tiles(tiles_m * tiles_n_), num_kb(num_kb_), tiles_n(tiles_n_), ragged(ragged_), round(0),
        worker(worker_), stride(workers), colmajor(colmajor_), rem_done(0) {}
  __device__ bool next(Seg& s) {
    s.atomic = 0;
    if (slices == 0) {
      const int v = round * stride + ((round & 1) ? stride - 1 - worker : worker);
      if (v >= num_tiles) return false;
      ++round;
      s.kb0 = 0; s.kb1 = num_kb;
      if (colmajor) {   // profiling knob: walk the tiles down the columns (ragged handling off)
        const int tm = num_tiles / tiles_n;
        s.tile = (v % tm) * tiles_n + v / tm; s.half = 0;
        return true;
      }
      const int n_full = tiles_n - ragged;
      const int count_full = (num_tiles / tiles_n) * n_full;
      if (v < count_full) { s.tile = (v / n_full) * tiles_n + (v % n_full); s.half = 0; }
      else { s.tile = (v - count_full) * tiles_n + tiles_n - 1; s.half = 1; }
      return true;
    }
    s.half = 0;
    const int full_rounds = num_tiles / stride;
    if (round < full_rounds) {
      s.tile = round * stride + ((round & 1) ? stride - 1 - worker : worker);
      ++round;
      s.kb0 = 0; s.kb1 = num_kb;
      return true;
    }
    const int rem = num_tiles - full_rounds * stride;
    if (rem_done || worker >= rem * slices) return false;
    rem_done = 1;
    const int sl = worker / rem;
    s.tile = full_rounds * stride + worker % rem;
    s.kb0 = (int)((int64_t)num_kb * sl / slices);
    s.kb1 = (int)((int64_t)num_kb * (sl + 1) / slices);
    s.atomic = slices > 1;
    return s.kb1 > s.kb0;
  }
};

// the ragged last column tile may run BN/2 wide when every CTA's share of it is still a whole number of TMA boxes
template <int CG, int BN, int B_MN>
__host__ __device__ constexpr bool half_tile_ok() {
  return B_MN ? ((BN / 2 / CG) % 64 == 0) : ((BN / 2 / CG) % 8 == 0 && (BN / 2) % 16 == 0);
}

// One epilogue warp's share of one accumulator tile: TMEM lane quadrant q (32 rows), every second 32-column chunk.
// Per chunk: tcgen05.ld (thread = row) -> smem transpose -> row-coalesced fused epilogue.  Everything a chunk needs
// from global memory (residual / saved pre-activation / old D rows, bias and gate vectors) is fetched in ONE batch
// of independent loads: for the first chunk before the accumulator is even complete (the latency hides behind the
// MMAs), for chunk i+1 right after the stores of chunk i.  No load is issued between a prefetch batch and its use,
// so a scoreboard wait never covers a younger load; the other seven epilogue warps fill the remaining latency.
template <int KIND, int BN, typename TD>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, TD* __restrict__ D, int64_t ldd, int M, int N, int m0,
                                              int n0, uint32_t taddr, float* __restrict__ st, int q, int half, int lane,
                                              uint64_t* tfull_bar, uint32_t tfull_phase) {
  constexpr bool kAuxF32 = KIND == kEpiGateRes || KIND == kEpiAccum;
  constexpr bool kAuxBf16 = KIND == kEpiDGelu || KIND == kEpiDSilu;
  constexpr int NCH = BN / kStageCols;
  const int rsub = lane >> 3, cc = (lane & 7) * 4;
  const int row_base = m0 + q * 32;

  const float* auxf = nullptr;
  const bf16* auxh = nullptr;
  int64_t ld_aux = 0;
  if constexpr (KIND == kEpiAccum) { auxf = reinterpret_cast<const float*>(D); ld_aux = ldd; }
  if constexpr (KIND == kEpiGateRes) { auxf = reinterpret_cast<const float*>(ep.aux); ld_aux = ep.ld_aux; }
  if constexpr (kAuxBf16) { auxh = reinterpret_cast<const bf16*>(ep.aux); ld_aux = ep.ld_aux; }

  float4 af[kAuxF32 ? 8 : 1];
  uint2 ah[kAuxBf16 ? 8 : 1];
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // gate rows: one per group of rows_per_group rows.  Normally a warp's 32 rows sit in one group (T % 32 == 0) and
  // the gate vector is fetched once per chunk; otherwise per row.
  int g_first = 0;
  bool g_uniform = true;
  if constexpr (KIND == kEpiGateRes) {
    g_first = row_base / ep.rows_per_group;
    const int last = (row_base + 31 < M ? row_base + 31 : M - 1) / ep.rows_per_group;
    g_uniform = g_first == last || row_base >= M;
  }

  const int n_lim = ep.bias_grad != nullptr ? ep.n_store : N;    // columns that exist in D
  auto prefetch = [&](int c0) {
    const int col = n0 + c0 + cc;
    if (col < n_lim) {
      if (ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
      if constexpr (KIND == kEpiGateRes) {
        if (g_uniform && row_base < M) g4 = __ldg(reinterpret_cast<const float4*>(ep.gate + (int64_t)g_first * ep.ld_gate + col));
      }
      if constexpr (kAuxF32 || kAuxBf16) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = row_base + it * 4 + rsub;
          if (row < M) {
            if constexpr (kAuxF32) af[it] = __ldg(reinterpret_cast<const float4*>(auxf + (int64_t)row * ld_aux + col));
            if constexpr (kAuxBf16) ah[it] = __ldg(reinterpret_cast<const uint2*>(auxh + (int64_t)row * ld_aux + col));
          }
        }
      }
    }
  };

  int ci = half;
  if (ci < NCH && n0 + ci * kStageCols < N) prefetch(ci * kStageCols);
  mbar_wait(tfull_bar, tfull_phase);
  tc_fence_after();

#pragma unroll 1
  for (; ci < NCH; ci += 2) {
    const int c0 = ci * kStageCols;
    if (n0 + c0 >= N) break;                 // warp-uniform
    {
      float v[32];
      tmem_ld32(taddr + c0, v);
      // thread = accumulator row: park the 32 columns in smem ...
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(st + lane * kStagePitch + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    // ... and pick them up row-coalesced: 8 lanes cover one 32-column row segment, 4 rows per instruction.
    // All eight shared-memory reads are issued as one batch, and a chunk that lies fully inside the matrix (the
    // common case, warp-uniform) runs without per-row predicates: with a branch per row the compiler serialises
    // LDS -> use eight times over.
    const int col = n0 + c0 + cc;
    float4 fr[8];
#pragma unroll
    for (int it = 0; it < 8; ++it)
      fr[it] = *reinterpret_cast<const float4*>(st + (it * 4 + rsub) * kStagePitch + (((lane & 7) ^ ((it * 4 + rsub) & 7)) << 2));
    const bool full = row_base + 32 <= M && n0 + c0 + kStageCols <= N;
    auto row_op = [&](int it) {
      const int row = row_base + it * 4 + rsub;
      const float4 f = fr[it];
      if constexpr (KIND == kEpiNone || KIND == kEpiAccum || KIND == kEpiAtomic) {
        if (ep.bias_grad != nullptr && col >= ep.n_store) {       // the ones column of B: column sums of A
          if (col == ep.n_store) atomicAdd(ep.bias_grad + row, f.x);
          return;
        }
      }
      F4 acc{{f.x + b4.x, f.y + b4.y, f.z + b4.z, f.w + b4.w}};
      TD* dptr = D + (int64_t)row * ldd + col;
      if constexpr (KIND == kEpiNone) {
        store4(dptr, acc);
      } else if constexpr (KIND == kEpiAccum) {
        const float4 o = af[it];
        acc.v[0] += o.x; acc.v[1] += o.y; acc.v[2] += o.z; acc.v[3] += o.w;
        store4(dptr, acc);
      } else if constexpr (KIND == kEpiAtomic) {
        red_add4(reinterpret_cast<float*>(dptr), acc);
      } else if constexpr (KIND == kEpiGelu || KIND == kEpiSilu) {
        if (ep.out2) store4(reinterpret_cast<bf16*>(ep.out2) + (int64_t)row * ep.ld_out2 + col, acc);
        F4 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float h = round_bf16(acc.v[i]);   // activate the value backward will see
          o.v[i] = KIND == kEpiGelu ? gelu_fast(h) : silu_fast(h);
        }
        store4(dptr, o);
      } else if constexpr (KIND == kEpiGateRes) {
        if (ep.out2) store4(reinterpret_cast<bf16*>(ep.out2) + (int64_t)row * ep.ld_out2 + col, acc);
        const float4 rs = af[it];
        float4 g = g4;
        if (!g_uniform) g = __ldg(reinterpret_cast<const float4*>(ep.gate + (int64_t)(row / ep.rows_per_group) * ep.ld_gate + col));
        F4 o{{rs.x + g.x * round_bf16(acc.v[0]), rs.y + g.y * round_bf16(acc.v[1]), rs.z + g.z * round_bf16(acc.v[2]),
              rs.w + g.w * round_bf16(acc.v[3])}};
        store4(dptr, o);
      } else {   // kEpiDGelu / kEpiDSilu
        const uint2 hv = ah[it];
        const __nv_bfloat162 h01 = *reinterpret_cast<const __nv_bfloat162*>(&hv.x);
        const __nv_bfloat162 h23 = *reinterpret_cast<const __nv_bfloat162*>(&hv.y);
        const float h[4] = {__low2float(h01), __high2float(h01), __low2float(h23), __high2float(h23)};
        F4 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) o.v[i] = acc.v[i] * (KIND == kEpiDGelu ? gelu_grad_fast(h[i]) : silu_grad_fast(h[i]));
        store4(dptr, o);
      }
    };
    if (full) {
#pragma unroll
      for (int it = 0; it < 8; ++it) row_op(it);
    } else if (col < N) {
#pragma unroll
      for (int it = 0; it < 8; ++it)
        if (row_base + it * 4 + rsub < M) row_op(it);
    }
    __syncwarp();
    if (ci + 2 < NCH && n0 + c0 + 2 * kStageCols < N) prefetch(c0 + 2 * kStageCols);
  }
}

// The epilogue's global operand of a tile (residual stream / saved pre-activation / old D rows): pulled into L2 one
// tile ahead by the warp that will consume it (lane = row, one 32-column segment per chunk), so the register
// prefetch inside epilogue_tile pays an L2 hit instead of an HBM round trip per chunk.
template <int BN>
__device__ __forceinline__ void epilogue_l2_prefetch(const char* aux, int64_t pitch_bytes, int esz, int M, int N, int m0,
                                                     int n0, int q, int half, int lane) {
  const int row = m0 + q * 32 + lane;
  if (aux == nullptr || row >= M) return;
  const char* rp = aux + (int64_t)row * pitch_bytes;
#pragma unroll
  for (int ci = half; ci < BN / kStageCols; ci += 2) {
    const int col = n0 + ci * kStageCols;
    if (col < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (int64_t)col * esz));
  }
}

// ------------------------------------------------------------------------------------------------
// TMA epilogue.  Every epilogue warp is its own pipeline over its chunks (32 accumulator rows x 32 columns), with no
// global loads or stores executed by its threads:
//   * the fused global operand of a chunk (saved pre-activation, bf16 / residual stream, fp32) arrives in the warp's
//     own shared-memory ring by TMA, requested 1-2 chunks ahead - across tile boundaries - and tracked by mbarriers,
//     not by register scoreboards;
//   * tcgen05.ld leaves thread = accumulator row; the math runs in that layout and reads the operand row from the
//     swizzled box (conflict-free 16-byte accesses);
//   * results are packed into a swizzled [32 x 32] box per output and written back by TMA stores (bulk groups; a box
//     is reused once the store of the previous chunk has read it - a second set of boxes would hide that wait, but
//     shared memory buys more as operand stages: these GEMMs lose 5-20 % going from 6 to 4 stages).
// bf16 outputs only (none / activation / activation-gradient).  Per-warp region: [operand slots x3][D box][out2 box]
// - 8 KB with a fused operand, 4 KB without.
// ------------------------------------------------------------------------------------------------
template <int KIND> struct TmaEpi {
  static constexpr bool kAuxBf16 = KIND == kEpiDGelu || KIND == kEpiDSilu;
  static constexpr bool kAuxF32 = false;                    // (the fp32 gate+residual epilogue stays on the register path)
  static constexpr bool kOut2 = KIND == kEpiGelu || KIND == kEpiSilu;
  static constexpr int kSlots = kAuxBf16 ? 3 : 0;
  static constexpr int kLook = 2;                           // operand requests in flight ahead of the chunk in work
  static_assert(kSlots == 0 || kSlots == kLook + 1, "a ring slot is reused by the request kLook + 1 chunks later");
  static constexpr int kSlotBytes = kAuxBf16 ? 2048 : 4096;
  static constexpr int kOutOff = kSlots * kSlotBytes;       // bf16 D box, 2 KB
  static constexpr int kOut2Off = kOutOff + 2048;
};

// 16-byte chunk j of row r inside a [32 x 64 B] SWIZZLE_64B box / a [32 x 128 B] SWIZZLE_128B box
__device__ __forceinline__ uint32_t sw64(uint32_t base, int r, int j) { return base + r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }
__device__ __forceinline__ uint32_t sw128(uint32_t base, int r, int j) { return base + r * 128 + ((j ^ (r & 7)) << 4); }
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int KIND, int CG, int BN, typename TD>
__device__ __forceinline__ void epilogue_loop_tma(const EpiParams& ep, const EpiMaps& em, int M, int N, int tiles_n,
                                                  uint32_t rank, Sched sched, uint32_t tmem_base, uint32_t wbuf,
                                                  uint64_t* auxbar, uint64_t* tfull, uint64_t* tempty, int q, int half,
                                                  int lane) {
  using E = TmaEpi<KIND>;
  constexpr int BMT = BM * CG, NCH = BN / kStageCols;
  static_assert(sizeof(TD) == 2, "TMA epilogue: bf16 D");
  // this warp's chunks of a tile that lie inside the matrix: chunk indices half, half + 2, ... (warp-uniform)
  auto my_chunks = [&](const Seg& sgm) {
    const int width = N - (sgm.tile % tiles_n) * BN;
    const int nvalid = width >= BN ? NCH : (width + kStageCols - 1) / kStageCols;
    return nvalid > half ? (nvalid - half + 1) / 2 : 0;
  };
  // tile -> first row of this warp / first column of the tile: one integer division per tile, not per chunk
  struct Org { int row0, n0; };
  auto org_of = [&](const Seg& sgm) {
    const int tm = sgm.tile / tiles_n;
    return Org{tm * BMT + (int)rank * BM + q * 32, (sgm.tile - tm * tiles_n) * BN};
  };
  auto col0_at = [&](const Org& o, int k) { return o.n0 + (half + 2 * k) * kStageCols; };

  int acc = 0;
  uint32_t acc_phase = 0;
  // bias of a chunk: lane l holds bias[col0 + l], fetched one chunk ahead (also across tiles), broadcast by shuffles
  auto load_bias = [&](const Org& o, int k) {
    const int c = col0_at(o, k) + lane;
    return (ep.bias != nullptr && c < N) ? __ldg(ep.bias + c) : 0.f;
  };
  float b_next = 0.f;
  int gbase = 0;        // chunks this warp completed in earlier tiles (ring position of the tile's first chunk)
  int issued = 0;       // operand requests made, counted from the current tile's first chunk (may run into the next tile)
  Seg sg, nx;
  bool have = sched.next(sg);
  Org og = have ? org_of(sg) : Org{0, 0}, ogn = og;
  bool b_ready = false;   // b_next already holds the bias of the coming tile's first chunk
  while (have) {
    const bool have_next = sched.next(nx);
    if (have_next) ogn = org_of(nx);
    const int cnt = my_chunks(sg), cntn = have_next ? my_chunks(nx) : 0;
    // (a tile may hold no chunk of this warp - a last column tile at most 32 columns wide - so the look-ahead of the
    // previous tile's last chunk cannot be relied on)
    if (cnt > 0 && !b_ready) b_next = load_bias(og, 0);
    b_ready = false;
    const int row0 = og.row0;
    if constexpr (E::kSlots > 0) {   // the next tile's operand rows -> L2, so the TMA requests above hit there
      if (have_next)
        epilogue_l2_prefetch<BN>((const char*)ep.aux, ep.ld_aux * (E::kAuxF32 ? 4 : 2), E::kAuxF32 ? 4 : 2, M, N,
                                 ogn.row0 - q * 32, ogn.n0, q, half, lane);
    }
    const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
    // request operand boxes up to kLook chunks ahead of chunk p (lane 0 issues; mbarrier per ring slot)
    auto top_up = [&](int p) {
      if constexpr (E::kSlots > 0) {
        while (issued < p + E::kLook + 1 && issued < cnt + cntn) {
          const bool in_next = issued >= cnt;
          const Org& to = in_next ? ogn : og;
          const int k = in_next ? issued - cnt : issued;
          const int slot = (gbase + issued) % E::kSlots;
          if (lane == 0) {
            // in place (fp32): the slot still feeds the store of the chunk that used it last
            if constexpr (E::kAuxF32) bulk_wait_read<0>();
            const uint32_t bar = smem_u32(&auxbar[slot]);
            mbar_expect_tx_u32(bar, 32 * kStageCols * (E::kAuxF32 ? 4 : 2));
            tma_load_2d_u32(&em.aux, bar, wbuf + slot * E::kSlotBytes, col0_at(to, k), to.row0);
          }
          ++issued;
        }
      }
    };
    top_up(0);
    mbar_wait(&tfull[acc], acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int p = 0; p < cnt; ++p) {
      const int G = gbase + p;                      // ring position
      const int col0 = col0_at(og, p);
      const float b_cur = b_next;
      if (p + 1 < cnt) b_next = load_bias(og, p + 1);
      else if (cntn > 0) { b_next = load_bias(ogn, 0); b_ready = true; }
      float v[32];
      tmem_ld32(taddr + (half + 2 * p) * kStageCols, v);
      if (ep.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __shfl_sync(0xffffffffu, b_cur, i);
      }
      top_up(p);                    // keeps kLook operand boxes in flight (into the next tile at the end of this one)
      // the D / out2 boxes were handed to the store of the previous chunk
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      const uint32_t outb = wbuf + E::kOutOff, out2b = wbuf + E::kOut2Off;
      if constexpr (KIND == kEpiNone) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts_u4(sw64(outb, lane, j), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      } else if constexpr (KIND == kEpiGelu || KIND == kEpiSilu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t h[4], a[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i] = pack_bf16x2(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);     // the pre-activation backward will see
            const float x0 = bf16_lo(h[i]), x1 = bf16_hi(h[i]);
            a[i] = KIND == kEpiGelu ? pack_bf16x2(gelu_fast(x0), gelu_fast(x1)) : pack_bf16x2(silu_fast(x0), silu_fast(x1));
          }
          if (ep.out2) sts_u4(sw64(out2b, lane, j), h[0], h[1], h[2], h[3]);
          sts_u4(sw64(outb, lane, j), a[0], a[1], a[2], a[3]);
        }
      } else if constexpr (E::kAuxBf16) {
        const int slot = G % E::kSlots;
        mbar_wait(&auxbar[slot], (uint32_t)(G / E::kSlots) & 1u);
        const uint32_t ab = wbuf + slot * E::kSlotBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 hv = lds_u4(sw64(ab, lane, j));
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float h0 = bf16_lo(hw[i]), h1 = bf16_hi(hw[i]);
            const float d0 = KIND == kEpiDGelu ? gelu_grad_fast(h0) : silu_grad_fast(h0);
            const float d1 = KIND == kEpiDGelu ? gelu_grad_fast(h1) : silu_grad_fast(h1);
            o[i] = pack_bf16x2(v[8 * j + 2 * i] * d0, v[8 * j + 2 * i + 1] * d1);
          }
          sts_u4(sw64(outb, lane, j), o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async();          // this thread's box writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&em.d, outb, col0, row0);
        if constexpr (E::kOut2) {
          if (ep.out2) tma_store_2d(&em.o2, out2b, col0, row0);
        }
        bulk_commit();
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (CG == 1) mbar_arrive(&tempty[acc]);
      else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    gbase += cnt;
    issued -= cnt;
    sg = nx;
    og = ogn;
    have = have_next;
  }
  if (lane == 0) bulk_wait_all();   // the boxes must outlive their stores
  __syncwarp();
}

// TMA epilogue of the adaLN-Zero gate + residual GEMMs (attn.proj / mlp.fc2, sit.py:134-135): fp32 D = res + gate * bf16(y),
// y = acc + bias -> out2 (bf16).  Two things bound this epilogue.  (1) The TMA unit of an SM serves about one box row
// (<= 128 bytes) per 4 cycles, whatever the row length: 16-column boxes (64-byte rows, the first version) capped the
// epilogue at 3 TB/s chip-wide, so chunks are 32 columns - 128-byte fp32 rows (SWIZZLE_128B), 64-byte bf16 rows.
// (2) Residual bytes in flight: every warp owns a ring of R residual boxes ([32 x 32] fp32, 4 KB), the requests for the
// next R - 1 chunks outstanding while one is worked on, across tile boundaries (R is a launch parameter; the host passes 2).
// A box is updated IN PLACE (thread = row) and stored back by TMA; its slot is re-requested one chunk later, once the
// store has read it.  y leaves through one [32 x 32] bf16 box (2 KB).  Shared memory is what this epilogue competes for
// with the operand pipeline (fc2: 6 -> 3 stages cost 72 -> 89 us of main loop), hence the lean ring: 10 KB per warp.
// Bias and gate ride in one register each per chunk: lane l holds bias[col0 + l] / gate[g, col0 + l].
constexpr int kGateResSlotBytes = 4096;
constexpr int kGateResMaxSlots = 8;
__host__ __device__ constexpr int gateres_warp_bytes(int R) { return R * kGateResSlotBytes + 2048; }

template <int CG, int BN>
__device__ __forceinline__ void epilogue_loop_tma_gateres(const EpiParams& ep, const EpiMaps& em, int M, int N, int tiles_n,
                                                          uint32_t rank, Sched sched, uint32_t tmem_base, uint32_t wbuf,
                                                          uint64_t* auxbar, uint64_t* tfull, uint64_t* tempty, int q,
                                                          int half, int lane, int R) {
  constexpr int BMT = BM * CG, W = kStageCols, NCH = BN / W;
  const uint32_t out2_base = wbuf + R * kGateResSlotBytes;
  auto my_chunks = [&](const Seg& sgm) {
    const int width = N - (sgm.tile % tiles_n) * BN;
    const int nvalid = width >= BN ? NCH : (width + W - 1) / W;
    return nvalid > half ? (nvalid - half + 1) / 2 : 0;
  };
  auto row0_of = [&](const Seg& sgm) { return (sgm.tile / tiles_n) * BMT + (int)rank * BM + q * 32; };
  auto col0_of = [&](const Seg& sgm, int k) { return (sgm.tile % tiles_n) * BN + (half + 2 * k) * W; };
  auto load_bias = [&](const Seg& ts, int k) {
    const int c = col0_of(ts, k) + lane;
    return (ep.bias != nullptr && c < N) ? __ldg(ep.bias + c) : 0.f;
  };
  auto load_gate = [&](const Seg& ts, int k) {
    const int c = col0_of(ts, k) + lane;
    if (c >= N) return 0.f;
    const int r0 = row0_of(ts);
    return __ldg(ep.gate + (int64_t)((r0 < M ? r0 : M - 1) / ep.rows_per_group) * ep.ld_gate + c);
  };
  int acc = 0;
  uint32_t acc_phase = 0;
  float b_next = 0.f, g_next = 0.f;
  int issued = 0;                      // residual requests made, counted from the current tile's first chunk
  int ld_slot = 0;                     // ring slot of the next request
  int use_slot = 0;                    // ring slot of the chunk in work ...
  uint32_t use_phase = 0;              // ... its mbarrier parity
  Seg sg, nx;
  bool have = sched.next(sg);
  bool bg_ready = false;
  while (have) {
    const bool have_next = sched.next(nx);
    const int cnt = my_chunks(sg), cntn = have_next ? my_chunks(nx) : 0;
    if (cnt > 0 && !bg_ready) { b_next = load_bias(sg, 0); g_next = load_gate(sg, 0); }
    bg_ready = false;
    const int row0 = row0_of(sg);
    const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
    // keep requests for the chunks p .. p + R - 1 outstanding (running into the next tile at the end of this one).  The
    // caller has made sure the store of chunk p - 1 has read its slot and the y box (wait_group.read 0).
    auto top_up = [&](int p) {
      while (issued < p + R && issued < cnt + cntn) {
        const bool in_next = issued >= cnt;
        const Seg& ts = in_next ? nx : sg;
        const int k = in_next ? issued - cnt : issued;
        if (lane == 0) {
          const uint32_t bar = smem_u32(&auxbar[ld_slot]);
          mbar_expect_tx_u32(bar, 32 * W * 4);
          tma_load_2d_u32(&em.aux, bar, wbuf + ld_slot * kGateResSlotBytes, col0_of(ts, k), row0_of(ts));
        }
        ++issued;
        if (++ld_slot == R) ld_slot = 0;
      }
    };
    if (lane == 0) bulk_wait_read<0>();
    top_up(0);
    mbar_wait(&tfull[acc], acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int p = 0; p < cnt; ++p) {
      const int col0 = col0_of(sg, p);
      const float b_cur = b_next, g_cur = g_next;
      if (p + 1 < cnt) { b_next = load_bias(sg, p + 1); g_next = load_gate(sg, p + 1); }
      else if (cntn > 0) { b_next = load_bias(nx, 0); g_next = load_gate(nx, 0); bg_ready = true; }
      const uint32_t ab = wbuf + use_slot * kGateResSlotBytes, out2b = out2_base;
      uint32_t y[16];
      {
        float v[32];
        tmem_ld32(taddr + (half + 2 * p) * W, v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          y[i] = pack_bf16x2(v[2 * i] + __shfl_sync(0xffffffffu, b_cur, 2 * i), v[2 * i + 1] + __shfl_sync(0xffffffffu, b_cur, 2 * i + 1));
      }
      // the stores of the previous chunk have read the y box and their residual slot: the slot takes the request for
      // chunk p + R - 1 now, while this chunk is still being worked on
      if (lane == 0) bulk_wait_read<0>();
      top_up(p);
      __syncwarp();
      if (ep.out2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_u4(sw64(out2b, lane, j), y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
      }
      mbar_wait(&auxbar[use_slot], use_phase);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t ra = sw128(ab, lane, j);
        const uint4 rv = lds_u4(ra);
        const float g0 = __shfl_sync(0xffffffffu, g_cur, 4 * j), g1 = __shfl_sync(0xffffffffu, g_cur, 4 * j + 1);
        const float g2 = __shfl_sync(0xffffffffu, g_cur, 4 * j + 2), g3 = __shfl_sync(0xffffffffu, g_cur, 4 * j + 3);
        const float o0 = __uint_as_float(rv.x) + g0 * bf16_lo(y[2 * j]);
        const float o1 = __uint_as_float(rv.y) + g1 * bf16_hi(y[2 * j]);
        const float o2 = __uint_as_float(rv.z) + g2 * bf16_lo(y[2 * j + 1]);
        const float o3 = __uint_as_float(rv.w) + g3 * bf16_hi(y[2 * j + 1]);
        sts_u4(ra, __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&em.d, ab, col0, row0);
        if (ep.out2) tma_store_2d(&em.o2, out2b, col0, row0);
        bulk_commit();
      }
      if (++use_slot == R) { use_slot = 0; use_phase ^= 1u; }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (CG == 1) mbar_arrive(&tempty[acc]);
      else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    issued -= cnt;
    sg = nx;
    have = have_next;
  }
  if (lane == 0) bulk_wait_all();
  __syncwarp();
}

template <int CG, int BN, int A_MN, int B_MN, typename TD, typename GM = NoGroups>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ EpiMaps emaps, TD* __restrict__ D, int64_t ldd, int M, int N, int K,
                    EpiParams ep, int stream_k, int dbg, int stages, int epi_bytes, int tma_epi,
                    const __grid_constant__ GM gmaps) {
  constexpr bool kGrouped = sizeof(GM) > sizeof(NoGroups);
  // stages / epi_bytes: pipeline depth and size of the epilogue's shared-memory region (host: GemmCfg::stages_for);
  // tma_epi: 0 = register / LSU epilogue; 1 = TMA epilogue (epilogue_loop_tma; emaps valid); fp32 D (gate+residual):
  // the number of residual ring slots per warp (epilogue_loop_tma_gateres)
  // stream_k: 0 = data-parallel; n >= 1 = split mode with n k-slices for the tiles of the last partial round
  // dbg (profiling only, results are garbage): 1 = no TMA (MMA does not wait for operands), 2 = no MMA issue,
  // 4 = no epilogue work (accumulators released immediately); 8 = column-major tile order (results stay correct)
  using Cfg = GemmCfg<CG, BN>;
  const int S = stages;
  constexpr int BNL = Cfg::kBNL;
  constexpr int BMT = BM * CG;                 // output-tile rows of the CTA (pair)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, by pointer arithmetic on the __shared__ array so that the compiler keeps
  // the shared address space (ld.shared / st.shared for the epilogue staging, not generic accesses)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* epi_region = smem + S * Cfg::kStageBytes;      // 1024-byte aligned (kStageBytes is a multiple of 1024)
  float* staging = reinterpret_cast<float*>(epi_region);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_region + epi_bytes);
  uint64_t* full = bars;              // [<=8] TMA bytes landed (CG = 2: the leader's copy counts both CTAs' bytes)
  uint64_t* empty = bars + 8;         // [<=8] MMAs reading the stage retired (arrives in every CTA of the pair)
  uint64_t* tfull = bars + 16;        // [2]   accumulator complete (arrives in every CTA of the pair)
  uint64_t* tempty = bars + 18;       // [2]   accumulator drained by all epilogue warps (of both CTAs; leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t* auxbars = bars + 24;      // [8 warps][8] operand boxes of the TMA epilogue landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
  const bool leader = rank == 0;
  const int worker = blockIdx.x / CG, workers = gridDim.x / CG;
  const int tiles_m = (M + BMT - 1) / BMT, tiles_n = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int n_rem = N - (tiles_n - 1) * BN;    // width of the last column tile
  const int ragged = (!stream_k && !(dbg & 8) && half_tile_ok<CG, BN, B_MN>() && n_rem <= BN / 2) ? 1 : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (tma_epi) {
      tma_prefetch_desc(&emaps.d);
      if (ep.out2 != nullptr) tma_prefetch_desc(&emaps.o2);
      if (ep.aux != nullptr) tma_prefetch_desc(&emaps.aux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiWarps * CG);
    }
    for (int s = 0; s < kEpiWarps * 8; ++s) mbar_init(&auxbars[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barriers, TMEM, descriptor prefetch) touched no global data: under programmatic dependent launch
  // it ran while the previous kernel of the stream was still draining.  From here on its results are needed.
  pdl_launch();
  pdl_wait();

  // The TMA and MMA warps run warp-uniform code: all 32 lanes walk the schedule and poll the barriers, one elected
  // lane executes the TMA / tcgen05 instructions.  Those take their operands from uniform registers; issued from a
  // single-lane divergent branch every one of them is wrapped in an R2UR "waterfall" loop of ~25 dependent
  // instructions, which costs more than the 64..128 cycles an MMA occupies the tensor pipe.
  if (warp == 0) {
    // ============================== TMA producer (every CTA loads its own A rows / B rows) ==============================
    const bool issuer = elect_one_sync();
    int stage = 0;
    uint32_t phase = 0;
    Sched sched(stream_k, tiles_m, tiles_n, ragged, num_kb, worker, workers, (dbg >> 3) & 1);
    Seg sg;
    const uint32_t stage0 = smem_u32(stage_base);
    const uint32_t full0 = smem_u32(full);
    while (!(dbg & 1) && sched.next(sg)) {
      const int m0 = (sg.tile / tiles_n) * BMT + (int)rank * BM;
      // a half tile is BN/2 wide: each CTA of the pair supplies BN/2/CG rows of B, taken from the head of its box
      int n0 = (sg.tile % tiles_n) * BN + (int)rank * (sg.half ? BNL / 2 : BNL);
      const CUtensorMap* pa = &map_a;
      const CUtensorMap* pb = &map_b;
      if constexpr (kGrouped) {
        if (gmaps.mode == 0) {                 // the column tile's group supplies B
          const int g = n0 / gmaps.per_group;
          n0 -= g * gmaps.per_group;
          pb = &gmaps.b[g];
        }
      }
      for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t sa = stage0 + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kABytes;
        const uint32_t fbar = full0 + stage * 8;
        int k0 = kb * BK;
        if constexpr (kGrouped) {
          if (gmaps.mode == 1) {               // the k-block's group supplies both operands
            const int g = kb / gmaps.per_group;
            k0 = (kb - g * gmaps.per_group) * BK;
            pa = &gmaps.a[g];
            pb = &gmaps.b[g];
          }
        }
        if (issuer) {
          if constexpr (CG == 1) {
            mbar_expect_tx_u32(fbar, Cfg::kStageBytes);
            if (A_MN) {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d_u32(pa, fbar, sa + c * (BK * 128), m0 + c * 64, k0);
            } else {
              tma_load_2d_u32(pa, fbar, sa, k0, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_2d_u32(pb, fbar, sb + c * (BK * 128), n0 + c * 64, k0);
            } else {
              tma_load_2d_u32(pb, fbar, sb, k0, n0);
            }
          } else {
            static_assert(!kGrouped || CG == 1, "grouped operands: cta_group::1 kernels only");
            if (leader) mbar_expect_tx_u32(fbar, CG * Cfg::kStageBytes);
            const uint32_t fb = mapa_u32(fbar, 0);
            if (A_MN) {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d_cg2_u32(&map_a, fb, sa + c * (BK * 128), m0 + c * 64, k0);
            } else {
              tma_load_2d_cg2_u32(&map_a, fb, sa, k0, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_2d_cg2_u32(&map_b, fb, sb + c * (BK * 128), n0 + c * 64, k0);
            } else {
              tma_load_2d_cg2_u32(&map_b, fb, sb, k0, n0);
            }
          }
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ============================== MMA issuer (leader CTA only) ==============================
    constexpr uint32_t idesc_full = make_idesc(BMT, BN, A_MN, B_MN);
    constexpr uint32_t idesc_half = make_idesc(BMT, BN / 2, A_MN, B_MN);
    // descriptor halves (cute::UMMA::SmemDescriptor): lo = start address >> 4 | LBO >> 4 << 16, hi = SBO >> 4 | version
    // | SWIZZLE_128B.  K-major: +32 B per 16-element k step inside the swizzled 128 B row (LBO unused, encoded 1);
    // MN-major: +16 k-rows of 128 B per k step, LBO = BK * 128 B to the next 64-element mn box.
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kLboA = (A_MN ? (uint32_t)(BK * 128) >> 4 : 1u) << 16, kStepA = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    constexpr uint32_t kLboB = (B_MN ? (uint32_t)(BK * 128) >> 4 : 1u) << 16, kStepB = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    const bool issuer = elect_one_sync();
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    Sched sched(stream_k, tiles_m, tiles_n, ragged, num_kb, worker, workers, (dbg >> 3) & 1);
    Seg sg;
    const uint32_t stage0 = smem_u32(stage_base);
    while (sched.next(sg)) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      const uint32_t idesc = sg.half ? idesc_half : idesc_full;
      for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
        if (!(dbg & 1)) mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = stage0 + stage * Cfg::kStageBytes;
        const uint32_t la = ((sa >> 4) & 0x3FFFu) | kLboA;
        const uint32_t lb = (((sa + Cfg::kABytes) >> 4) & 0x3FFFu) | kLboB;
        if (issuer) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = ((uint64_t)kDescHi << 32) | (la + k * kStepA);
            const uint64_t db = ((uint64_t)kDescHi << 32) | (lb + k * kStepB);
            if (!(dbg & 2)) umma_bf16<CG>(tmem_d, da, db, idesc, (kb > sg.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit<CG>(&empty[stage]);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (issuer) umma_commit<CG>(&tfull[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ============================== epilogue (8 warps per CTA, this CTA's 128 accumulator rows) ==============================
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int half = (warp - kFirstEpiWarp) >> 2;            // which 32-column chunks (even / odd) this warp drains
    float* st = staging + (warp - kFirstEpiWarp) * 32 * kStagePitch;
    int acc = 0;
    uint32_t acc_phase = 0;
    Sched sched(stream_k, tiles_m, tiles_n, ragged, num_kb, worker, workers, (dbg >> 3) & 1);
    if (tma_epi && !(dbg & 4)) {
      const uint32_t wbuf = smem_u32(epi_region) + (warp - kFirstEpiWarp) * (epi_bytes / kEpiWarps);
      uint64_t* ab = auxbars + (warp - kFirstEpiWarp) * 8;
#define REED_TMA_EPI(KIND) epilogue_loop_tma<KIND, CG, BN, TD>(ep, emaps, M, N, tiles_n, rank, sched, tmem_base, wbuf, ab, tfull, tempty, q, half, lane)
      if constexpr (sizeof(TD) == 4) {
        epilogue_loop_tma_gateres<CG, BN>(ep, emaps, M, N, tiles_n, rank, sched, tmem_base, wbuf, ab, tfull, tempty, q, half, lane, tma_epi);
      } else {
        if (ep.kind == kEpiNone) REED_TMA_EPI(kEpiNone);
        else if (ep.kind == kEpiGelu) REED_TMA_EPI(kEpiGelu);
        else if (ep.kind == kEpiSilu) REED_TMA_EPI(kEpiSilu);
        else if (ep.kind == kEpiDGelu) REED_TMA_EPI(kEpiDGelu);
        else REED_TMA_EPI(kEpiDSilu);
      }
#undef REED_TMA_EPI
    } else {
    // global operand of the fused epilogue, if any (see epilogue_l2_prefetch)
    const char* aux = nullptr;
    int64_t aux_pitch = 0;
    int aux_esz = 4;
    if (ep.kind == kEpiGateRes) { aux = (const char*)ep.aux; aux_pitch = ep.ld_aux * 4; }
    else if (ep.kind == kEpiDGelu || ep.kind == kEpiDSilu) { aux = (const char*)ep.aux; aux_pitch = ep.ld_aux * 2; aux_esz = 2; }
    else if (ep.kind == kEpiNone && ep.accumulate && sizeof(TD) == 4) { aux = (const char*)D; aux_pitch = ldd * 4; }
    const int n_cols = ep.bias_grad != nullptr ? ep.n_store : N;   // columns that exist in D
    Seg sg, nx;
    bool have = sched.next(sg);
    if (have && !sg.atomic)
      epilogue_l2_prefetch<BN>(aux, aux_pitch, aux_esz, M, n_cols, (sg.tile / tiles_n) * BMT + (int)rank * BM, (sg.tile % tiles_n) * BN, q, half, lane);
    while (have) {
      const bool have_next = sched.next(nx);
      if (have_next && !nx.atomic)
        epilogue_l2_prefetch<BN>(aux, aux_pitch, aux_esz, M, n_cols, (nx.tile / tiles_n) * BMT + (int)rank * BM, (nx.tile % tiles_n) * BN, q, half, lane);
      const int m0 = (sg.tile / tiles_n) * BMT + (int)rank * BM, n0 = (sg.tile % tiles_n) * BN;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
#define REED_EPI(KIND) epilogue_tile<KIND, BN, TD>(ep, D, ldd, M, N, m0, n0, taddr, st, q, half, lane, &tfull[acc], acc_phase)
      if (dbg & 4) {
        mbar_wait(&tfull[acc], acc_phase);
      } else if constexpr (sizeof(TD) == 4) {
        if (sg.atomic) REED_EPI(kEpiAtomic);
        else if (ep.kind == kEpiGateRes) REED_EPI(kEpiGateRes);
        else if (ep.kind == kEpiNone && ep.accumulate) REED_EPI(kEpiAccum);
        else if (ep.kind == kEpiNone) REED_EPI(kEpiNone);
        else if (ep.kind == kEpiGelu) REED_EPI(kEpiGelu);
        else if (ep.kind == kEpiSilu) REED_EPI(kEpiSilu);
        else if (ep.kind == kEpiDGelu) REED_EPI(kEpiDGelu);
        else REED_EPI(kEpiDSilu);
      } else {
        if (ep.kind == kEpiNone) REED_EPI(kEpiNone);
        else if (ep.kind == kEpiGelu) REED_EPI(kEpiGelu);
        else if (ep.kind == kEpiSilu) REED_EPI(kEpiSilu);
        else if (ep.kind == kEpiDGelu) REED_EPI(kEpiDGelu);
        else REED_EPI(kEpiDSilu);
      }
#undef REED_EPI
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tempty[acc]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      sg = nx;
      have = have_next;
    }
    }
  }
  tc_fence_before();
  __syncwarp();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // nobody leaves while the peer may still signal it or read its smem
  if (warp == 2) tmem_dealloc<CG>(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// launch helpers (instantiated in gemm_tcgen05_cg1.cu / gemm_tcgen05_cg2.cu)
// ------------------------------------------------------------------------------------------------
template <int CG, int BN, int A_MN, int B_MN, typename TD, typename GM = NoGroups>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd, int M, int N, int K,
                  const EpiParams& ep, cudaStream_t st, int grid, int stream_k, const EpiMaps* em, const GM* gm = nullptr) {
  static const int dbg = getenv("REED_GEMM_DEBUG") ? atoi(getenv("REED_GEMM_DEBUG")) : 0;
  using Cfg = GemmCfg<CG, BN>;
  static_assert(Cfg::stages_for(Cfg::kStagingBytes) >= 3, "pipeline too shallow");
  static_assert(!B_MN || Cfg::kBNL % 64 == 0, "MN-major B is staged in 64-column TMA boxes");
  constexpr bool kTmaFits = Cfg::stages_for(Cfg::kTmaEpiAuxBytes) >= 3;
  int tma_epi = (em != nullptr && kTmaFits) ? 1 : 0;
  const bool fused_operand = ep.kind == kEpiDGelu || ep.kind == kEpiDSilu;
  int epi_bytes = tma_epi ? (fused_operand ? Cfg::kTmaEpiAuxBytes : Cfg::kTmaEpiBytes) : Cfg::kStagingBytes;
  if (tma_epi && sizeof(TD) == 4) {
    // gate+residual: two residual ring slots per epilogue warp (the ring competes with the operand stages for shared
    // memory; deeper rings measured no faster, profiles/r02_gemm_notes.md); tile configurations whose stages are too
    // large even for that keep the register epilogue
    constexpr int R = 2;
    if (Cfg::stages_for(kEpiWarps * gateres_warp_bytes(R)) >= 3) {
      tma_epi = R;
      epi_bytes = kEpiWarps * gateres_warp_bytes(R);
    } else {
      tma_epi = 0;
      epi_bytes = Cfg::kStagingBytes;
    }
  }
  static const int max_stages = getenv("REED_GEMM_MAX_STAGES") ? atoi(getenv("REED_GEMM_MAX_STAGES")) : 8;   // profiling knob
  const int stages = Cfg::stages_for(epi_bytes) < max_stages ? Cfg::stages_for(epi_bytes) : (max_stages < 2 ? 2 : max_stages);
  const int smem = Cfg::smem_bytes(stages, epi_bytes);
  auto kernel = gemm_tcgen05_kernel<CG, BN, A_MN, B_MN, TD, GM>;
  static bool configured = false;   // per template instance
  if (!configured) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  static const int pdl = getenv("REED_PDL") ? atoi(getenv("REED_PDL")) : 1;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  static const EpiMaps no_maps{};
  static const GM no_groups{};
  REED_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, ma, mb, tma_epi ? *em : no_maps, (TD*)D, ldd, M, N, K, ep, stream_k, dbg,
                                     stages, epi_bytes, tma_epi, gm != nullptr ? *gm : no_groups));
  return 0;
}

template <int CG, int BN, typename TD>
static int launch_major(int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd, int M,
                        int N, int K, const EpiParams& ep, cudaStream_t st, int grid, int stream_k, const EpiMaps* em) {
  if (!a_mn && !b_mn) return launch<CG, BN, 0, 0, TD>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, em);
  if constexpr ((BN / CG) % 64 == 0) {
    if (!a_mn && b_mn) return launch<CG, BN, 0, 1, TD>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, em);
    if (a_mn && b_mn) return launch<CG, BN, 1, 1, TD>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, em);
  }
  if (a_mn && !b_mn) return launch<CG, BN, 1, 0, TD>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, em);
  return fail("gemm_tcgen05: no kernel for cta_group::%d BN=%d with MN-major B", CG, BN);
}

template <int CG>
static int launch_cg(int bn, int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd,
                     int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st, int grid, int stream_k,
                     const EpiMaps* em) {
#define GO(BNV)                                                                                                        \
  (d_dtype == kF32 ? launch_major<CG, BNV, float>(a_mn, b_mn, ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, em)     \
                   : launch_major<CG, BNV, bf16>(a_mn, b_mn, ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, em))
  if (bn == 256) return GO(256);
  if (bn == 192) return GO(192);
  return GO(128);
#undef GO
}

}  // namespace reed
