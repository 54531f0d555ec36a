// Building blocks shared by the tcgen05 attention kernels (attention_fa.cu forward, attention_fa_bwd.cu backward):
// operand tiles as TMA stages them, shared-memory / tensor-memory MMA forms, TMEM load/store shapes, tensor maps.
#pragma once
#include <cuda.h>
#include <math.h>
#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace reed {
namespace fa {

constexpr int kRows = 128;            // rows of every operand tile (queries or keys per block)

// head_dim 72 is not a multiple of the 64-element swizzle row: every [128 x hd] operand tile is staged as a [128 x 64]
// SWIZZLE_128B tile plus a [128 x 16] SWIZZLE_32B tail whose columns 72..79 are zero-filled by TMA (the tensor map's
// innermost extent is hd, so they are out of bounds).  Contractions over head_dim take 4 + 1 k-steps; outputs over
// head_dim are two MMAs (N = 64 and N = 16) into adjacent TMEM columns.
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
// MN-major operand spanning more than 64 mn-elements (two [rows x 64] SWIZZLE_128B tiles `lbo_bytes` apart)
__device__ __forceinline__ uint32_t desc_lo_lbo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
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

// tcgen05.mma with the A operand in tensor memory: bf16 A[128 x K] = lane per row, 32-bit column c = elements (2c, 2c+1),
// 8 columns per K = 16 step (layout verified on the B200 by profiles/probe_ts_mma.cu)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
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
// D[128 x hd] (+)= P[128 x 128] . Z[128 x hd]: P = two K-major SWIZZLE_128B tiles in shared memory,
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
// Same product with P in TENSOR MEMORY (64 columns of bf16 pairs at tmem_p, written by the softmax warps)
// `ones` (optional): a [16 x 16] SWIZZLE_32B tile whose column 0 is 1 -> column kND of D accumulates the row sums of P
template <int HD>
__device__ __forceinline__ void mma_pv_ts(bool leader, uint32_t tmem_d, uint32_t tmem_p, uint32_t tile_z, bool accumulate,
                                          uint32_t ones = 0) {
  constexpr uint32_t idesc64 = make_idesc(128, 64, 0, 1);
  constexpr uint32_t idesc16 = make_idesc(128, 16, 0, 1);
  const uint32_t lz = desc_lo(tile_z);
  const uint64_t d_ones = mk_desc(desc_lo(ones), kHiSw32);
  if (leader) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t acc = (accumulate || ks > 0) ? 1u : 0u;
      umma_ts(tmem_d, tmem_p + ks * 8, mk_desc(lz + ks * (2048 >> 4), kHiSw128), idesc64, acc);
      if (Tile<HD>::kTail)
        umma_ts(tmem_d + 64, tmem_p + ks * 8, mk_desc(lz + (Tile<HD>::kMain >> 4) + ks * (512 >> 4), kHiSw32), idesc16, acc);
      if (ones != 0) umma_ts(tmem_d + Tile<HD>::kND, tmem_p + ks * 8, d_ones, idesc16, acc);
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
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// one [128 x hd] tile: rows row0.., "head" column block hcol of the packed tensor
template <int HD>
__device__ __forceinline__ void load_tile(const CUtensorMap* main, const CUtensorMap* tail, uint64_t* bar, uint32_t dst,
                                          int hcol, int row0) {
  tma_load_3d(main, bar, dst, 0, hcol, row0);
  if (Tile<HD>::kTail) tma_load_3d(tail, bar, dst + Tile<HD>::kMain, 64, hcol, row0);
}
// pull the same tile into L2 only (no shared memory, no barrier): issued a step ahead of the load that needs it
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
template <int HD>
__device__ __forceinline__ void prefetch_tile(const CUtensorMap* main, const CUtensorMap* tail, int hcol, int row0) {
  tma_prefetch_3d(main, 0, hcol, row0);
  if (Tile<HD>::kTail) tma_prefetch_3d(tail, 64, hcol, row0);
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
// 16 consecutive TMEM columns of this thread's lane <- registers (pair with tmem_wait_st)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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
  CUtensorMap o_main, o_tail;       // [B*T, H, hd]   saved context (backward: delta = rowsum(dO * O))
  CUtensorMap do_main, do_tail;     // [B*T, H, hd]
  CUtensorMap out_main, out_tail8;  // stores: forward -> o, backward -> dqkv (tail8 = 8-column box, no swizzle)
};

// ------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
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
inline int make_map3(CUtensorMap* map, const void* ptr, int64_t rows, int heads, int hd, int kind) {
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

inline int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = kNumSMs;
  }
  return sms;
}

}  // namespace fa
}  // namespace reed
