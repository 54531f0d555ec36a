// Fused attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA),
// forward and backward, for the 256-token patch sequence of the SiT hot path (T = 128 or 256; head_dim 64 or 72).
// Reads Q/K/V straight out of the packed qkv GEMM output [B,T,3,H,hd] through 3-D tensor maps, writes the context
// [B,T,H,hd] and dqkv in the packed layout.  Longer sequences (T = 1024) run the mma.sync kernels of attention_mma.cu.
//
// With T <= 256 the whole score row of a query lives in TMEM (128 lanes x 256 fp32 columns), so there is no online
// softmax and no accumulator rescaling: S = Q K^T is issued for both 128-key blocks up front, the four softmax warps
// (thread = query row = TMEM lane) take the row max over all columns, then exponentiate block by block, writing
// bf16 P into shared memory in the K-major SWIZZLE_128B layout the next MMA (O += P V) consumes.
//
// head_dim 72 is not a multiple of the 64-element swizzle row: every [128 x hd] operand tile is staged as a
// [128 x 64] SWIZZLE_128B tile plus a [128 x 16] SWIZZLE_32B tail whose columns 72..79 are zero-filled by TMA
// (the tensor map's innermost extent is hd, so they are out of bounds).  Contractions over head_dim take 4 + 1
// k-steps; outputs over head_dim are two MMAs (N = 64 and N = 16) into adjacent TMEM columns.
//
//   forward : CTA = 128 query rows of one (batch, head).                 TMEM: S0 | S1 (O aliases S0 once P0 is out)
//   dQ      : CTA = 128 query rows; S_j = Q K_j^T, dP_j = dO V_j^T, dS_j = P_j (dP_j - delta) -> smem, dQ += dS_j K_j
//             (also emits delta = rowsum(dO * O)).                       TMEM: S0 dP0 S1 dP1 (dQ aliases S0)
//   dK/dV   : CTA = 128 key rows, everything transposed (lanes = keys): S^T_i = K Q_i^T, dP^T_i = V dO_i^T,
//             dV += P^T_i dO_i, dK += dS^T_i Q_i.                        TMEM: S^T0 dP^T0 S^T1 dP^T1 (dV, dK alias block 0)
//
// Reference semantics: timm Attention.forward with fused_attn (F.scaled_dot_product_attention, scale hd^-0.5),
// imported at /root/reference/image/models/sit.py:13 and called at sit.py:134; backward = autograd of the same.
#include <cuda.h>
#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace reed {

namespace {

constexpr int kRows = 128;            // rows of every operand tile (queries or keys per block)
constexpr int kThreads = 160;         // warps 0-3: softmax / elementwise (TMEM lane quadrants 0-3); warp 4: TMA + MMA
constexpr uint32_t kSw128 = 2, kSw32 = 6;

template <int HD> struct Tile {
  static constexpr bool kTail = HD > 64;
  static constexpr int kMain = kRows * 128;                 // [128 x 64] bf16, SWIZZLE_128B
  static constexpr int kTailBytes = kTail ? kRows * 32 : 0; // [128 x 16] bf16, SWIZZLE_32B (cols 72..79 zero)
  static constexpr int kBytes = kMain + kTailBytes;         // 16384 / 20480: multiples of 1024
  static constexpr int kND = kTail ? 80 : 64;               // head_dim as the tensor core sees it
};
constexpr int kPBytes = 2 * kRows * 128;                    // [128 x 128] bf16 as two K-major SWIZZLE_128B tiles

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// K-major operand, k-step ks (16 elements) of a [128 x hd] tile staged as main + tail
template <int HD>
__device__ __forceinline__ uint64_t desc_k(uint32_t tile, int ks) {
  if (ks < 4) return smem_desc(tile + ks * 32, 16, 1024, kSw128);
  return smem_desc(tile + Tile<HD>::kMain, 16, 256, kSw32);
}
// K-major operand written by the softmax warps: [128 x 128] as two [128 x 64] SWIZZLE_128B tiles
__device__ __forceinline__ uint64_t desc_p(uint32_t base, int ks) {
  return smem_desc(base + (ks >> 2) * (kRows * 128) + (ks & 3) * 32, 16, 1024, kSw128);
}
// MN-major operand ([contraction rows x hd] tile read along its rows): k-step ks = 16 tile rows
__device__ __forceinline__ uint64_t desc_mn_main(uint32_t tile, int ks) { return smem_desc(tile + ks * 2048, 8192, 1024, kSw128); }
template <int HD>
__device__ __forceinline__ uint64_t desc_mn_tail(uint32_t tile, int ks) {
  return smem_desc(tile + Tile<HD>::kMain + ks * 512, 4096, 256, kSw32);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
  umma_bf16<1>(tmem_d, da, db, idesc, accumulate ? 1u : 0u);
}
__device__ __forceinline__ void commit(uint64_t* bar) { umma_commit<1>(bar); }

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// one [128 x hd] tile: rows row0.., "head" column block hcol of the packed tensor
template <int HD>
__device__ __forceinline__ void load_tile(const CUtensorMap* main, const CUtensorMap* tail, uint64_t* bar, uint32_t dst,
                                          int hcol, int row0) {
  tma_load_3d(main, bar, dst, 0, hcol, row0);
  if (Tile<HD>::kTail) tma_load_3d(tail, bar, dst + Tile<HD>::kMain, 64, hcol, row0);
}

// 32 consecutive TMEM columns of this thread's lane, no wait (pair with tmem_wait_ld)
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
// 32 values of row `row` (columns c0..c0+31 of a [128 x 128] K-major SWIZZLE_128B pair of tiles) -> shared memory.
// 16-byte chunk c of a row sits at chunk position c ^ (row & 7): the 8 lanes of a quarter-warp hit 8 distinct
// positions, so the stores are bank-conflict free.
__device__ __forceinline__ void store_p32(uint32_t base, int row, int c0, const float* v) {
  const uint32_t tile = base + (c0 >> 6) * (kRows * 128) + row * 128;
  const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t addr = tile + (((chunk0 + q) ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2(v[8 * q], v[8 * q + 1])),
                 "r"(pack2(v[8 * q + 2], v[8 * q + 3])), "r"(pack2(v[8 * q + 4], v[8 * q + 5])),
                 "r"(pack2(v[8 * q + 6], v[8 * q + 7]))
                 : "memory");
  }
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
// hd fp32 accumulator columns of this thread's TMEM lane -> scaled bf16 row in global memory
template <int HD>
__device__ __forceinline__ void store_acc_row(uint32_t taddr, float mul, bf16* dst) {
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    tmem_ld32_nowait(taddr + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 o;
      o.x = pack2(v[8 * q] * mul, v[8 * q + 1] * mul);
      o.y = pack2(v[8 * q + 2] * mul, v[8 * q + 3] * mul);
      o.z = pack2(v[8 * q + 4] * mul, v[8 * q + 5] * mul);
      o.w = pack2(v[8 * q + 6] * mul, v[8 * q + 7] * mul);
      *reinterpret_cast<uint4*>(dst + c0 + 8 * q) = o;
    }
  }
  if (Tile<HD>::kTail) {
    float v[16];
    tmem_ld16_nowait(taddr + 64, v);
    tmem_wait_ld();
    uint4 o;
    o.x = pack2(v[0] * mul, v[1] * mul);
    o.y = pack2(v[2] * mul, v[3] * mul);
    o.z = pack2(v[4] * mul, v[5] * mul);
    o.w = pack2(v[6] * mul, v[7] * mul);
    *reinterpret_cast<uint4*>(dst + 64) = o;
  }
}

struct AttnMaps {
  CUtensorMap qkv_main, qkv_tail;   // [B*T, 3H, hd]
  CUtensorMap o_main, o_tail;       // [B*T, H, hd]   (o in dq: the saved context; unused in forward)
  CUtensorMap do_main, do_tail;     // [B*T, H, hd]
};

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads, 2)
attn_tc5_fwd_kernel(const __grid_constant__ AttnMaps maps, bf16* __restrict__ o, float* __restrict__ lse, int T, int H,
                    float scale_log2) {
  using TL = Tile<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + TL::kBytes;            // 2 tiles; P aliases this region once both S blocks are complete
  const uint32_t sV = sK + 2 * TL::kBytes;        // 2 tiles
  const uint32_t sP = sK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TL::kBytes);
  uint64_t* bar_q = bars;          // Q landed
  uint64_t* bar_k = bars + 1;      // [2] K_j landed
  uint64_t* bar_v = bars + 3;      // [2] V_j landed
  uint64_t* bar_s = bars + 5;      // [2] S_j complete in TMEM
  uint64_t* bar_p = bars + 7;      // [2] P_j written by all 128 softmax threads
  uint64_t* bar_o = bars + 9;      // [2] O += P_j V_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nblk = T / kRows;
  const int row0 = b * T + qb * kRows;

  if (threadIdx.x == 128) {
    mbar_init(bar_q, 1);
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bar_k[j], 1);
      mbar_init(&bar_v[j], 1);
      mbar_init(&bar_s[j], 1);
      mbar_init(&bar_p[j], 128);
      mbar_init(&bar_o[j], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.qkv_main);
    if (TL::kTail) tma_prefetch_desc(&maps.qkv_tail);
  }
  if (warp == 4) tmem_alloc<1>(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA: everything this CTA will ever read, up front ----
      mbar_expect_tx(bar_q, TL::kBytes);
      load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bar_q, sQ, h, row0);
      for (int j = 0; j < nblk; ++j) {
        mbar_expect_tx(&bar_k[j], TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_k[j], sK + j * TL::kBytes, H + h, b * T + j * kRows);
      }
      for (int j = 0; j < nblk; ++j) {
        mbar_expect_tx(&bar_v[j], TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_v[j], sV + j * TL::kBytes, 2 * H + h, b * T + j * kRows);
      }
      // ---- S_j = Q K_j^T ----
      constexpr uint32_t idesc_s = make_idesc(128, 128, 0, 0);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&bar_k[j], 0);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TL::kND / 16; ++ks)
          umma(tmem + j * 128, desc_k<HD>(sQ, ks), desc_k<HD>(sK + j * TL::kBytes, ks), idesc_s, ks > 0);
        commit(&bar_s[j]);
      }
      // ---- O += P_j V_j  (O occupies columns 0..kND-1, over the dead S_0) ----
      constexpr uint32_t idesc_o64 = make_idesc(128, 64, 0, 1);
      constexpr uint32_t idesc_o16 = make_idesc(128, 16, 0, 1);
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&bar_v[j], 0);
        mbar_wait(&bar_p[j], 0);
        tc_fence_after();
        const uint32_t v = sV + j * TL::kBytes;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma(tmem, desc_p(sP, ks), desc_mn_main(v, ks), idesc_o64, j > 0 || ks > 0);
          if (TL::kTail) umma(tmem + 64, desc_p(sP, ks), desc_mn_tail<HD>(v, ks), idesc_o16, j > 0 || ks > 0);
        }
        commit(&bar_o[j]);
      }
    }
    __syncwarp();
  } else {
    // ---- softmax: thread = query row = TMEM lane ----
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    float m = -INFINITY;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&bar_s[j], 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 64) {
        float v[64];
        tmem_ld32_nowait(trow + j * 128 + c0, v);
        tmem_ld32_nowait(trow + j * 128 + c0 + 32, v + 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 64; ++i) m = fmaxf(m, v[i]);
      }
    }
    const float mc = m * scale_log2;
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      if (j > 0) mbar_wait(&bar_o[j - 1], 0);     // the MMA has finished reading P_{j-1}
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32_nowait(trow + j * 128 + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = ex2(fmaf(v[i], scale_log2, -mc));
          l += v[i];
        }
        store_p32(sP, row, c0, v);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar_p[j]);
    }
    mbar_wait(&bar_o[nblk - 1], 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int tok = qb * kRows + row;
    lse[((int64_t)b * H + h) * T + tok] = (mc + log2f(l)) * 0.6931471805599453f;
    store_acc_row<HD>(trow, inv, o + ((int64_t)b * T + tok) * ((int64_t)H * HD) + h * HD);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<1>(tmem, 256);
}

// ------------------------------------------------------------------------------------------------------
// backward, part 1: dQ and delta.  CTA = 128 query rows.
// ------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc5_dq_kernel(const __grid_constant__ AttnMaps maps, const float* __restrict__ lse, bf16* __restrict__ dqkv,
                   float* __restrict__ delta, int T, int H, float scale, float scale_log2) {
  using TL = Tile<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sDO = sQ + TL::kBytes;
  const uint32_t sK = sDO + TL::kBytes;           // 2 tiles
  const uint32_t sV = sK + 2 * TL::kBytes;        // 2 tiles
  const uint32_t sDS = sV + 2 * TL::kBytes;       // kPBytes; the saved context O is staged here first (for delta)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TL::kBytes + kPBytes);
  uint64_t* bar_q = bars;          // Q, dO, O landed
  uint64_t* bar_kv = bars + 1;     // [2] K_j, V_j landed
  uint64_t* bar_s = bars + 3;      // [2] S_j and dP_j complete
  uint64_t* bar_ds = bars + 5;     // [2] dS_j written
  uint64_t* bar_dq = bars + 7;     // [2] dQ += dS_j K_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nblk = T / kRows;
  const int row0 = b * T + qb * kRows;

  if (threadIdx.x == 128) {
    mbar_init(bar_q, 1);
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bar_kv[j], 1);
      mbar_init(&bar_s[j], 1);
      mbar_init(&bar_ds[j], 128);
      mbar_init(&bar_dq[j], 1);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, 3 * TL::kBytes);
      load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bar_q, sQ, h, row0);
      load_tile<HD>(&maps.do_main, &maps.do_tail, bar_q, sDO, h, row0);
      load_tile<HD>(&maps.o_main, &maps.o_tail, bar_q, sDS, h, row0);
      for (int j = 0; j < nblk; ++j) {
        mbar_expect_tx(&bar_kv[j], 2 * TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_kv[j], sK + j * TL::kBytes, H + h, b * T + j * kRows);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_kv[j], sV + j * TL::kBytes, 2 * H + h, b * T + j * kRows);
      }
      constexpr uint32_t idesc_s = make_idesc(128, 128, 0, 0);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&bar_kv[j], 0);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TL::kND / 16; ++ks)     // S_j = Q K_j^T           -> columns j*256 ..
          umma(tmem + j * 256, desc_k<HD>(sQ, ks), desc_k<HD>(sK + j * TL::kBytes, ks), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < TL::kND / 16; ++ks)     // dP_j = dO V_j^T         -> columns j*256 + 128 ..
          umma(tmem + j * 256 + 128, desc_k<HD>(sDO, ks), desc_k<HD>(sV + j * TL::kBytes, ks), idesc_s, ks > 0);
        commit(&bar_s[j]);
      }
      constexpr uint32_t idesc_o64 = make_idesc(128, 64, 0, 1);
      constexpr uint32_t idesc_o16 = make_idesc(128, 16, 0, 1);
      for (int j = 0; j < nblk; ++j) {                // dQ += dS_j K_j  (columns 0..kND-1, over the dead S_0)
        mbar_wait(&bar_ds[j], 0);
        tc_fence_after();
        const uint32_t k = sK + j * TL::kBytes;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma(tmem, desc_p(sDS, ks), desc_mn_main(k, ks), idesc_o64, j > 0 || ks > 0);
          if (TL::kTail) umma(tmem + 64, desc_p(sDS, ks), desc_mn_tail<HD>(k, ks), idesc_o16, j > 0 || ks > 0);
        }
        commit(&bar_dq[j]);
      }
    }
    __syncwarp();
  } else {
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const int tok = qb * kRows + row;
    const int64_t sidx = ((int64_t)b * H + h) * T + tok;
    const float ls = lse[sidx] * 1.4426950408889634f;     // exp2 domain
    // delta = rowsum(dO * O) from the staged tiles
    mbar_wait(bar_q, 0);
    float dl = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) dl += dot8(load_tile_chunk<HD>(sDO, row, c), load_tile_chunk<HD>(sDS, row, c));
    delta[sidx] = dl;
    // every thread is done with the staged O before dS_0 overwrites it
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&bar_s[j], 0);
      tc_fence_after();
      if (j > 0) mbar_wait(&bar_dq[j - 1], 0);    // the MMA has finished reading dS_{j-1}
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float s[32], dp[32];
        tmem_ld32_nowait(trow + j * 256 + c0, s);
        tmem_ld32_nowait(trow + j * 256 + 128 + c0, dp);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = ex2(fmaf(s[i], scale_log2, -ls));
          s[i] = p * (dp[i] - dl);                 // dS without the softmax scale (applied to dQ at the end)
        }
        store_p32(sDS, row, c0, s);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar_ds[j]);
    }
    mbar_wait(&bar_dq[nblk - 1], 0);
    tc_fence_after();
    store_acc_row<HD>(trow, scale, dqkv + ((int64_t)b * T + tok) * (3LL * H * HD) + h * HD);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<1>(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------
// backward, part 2: dK and dV.  CTA = 128 key rows; lanes = keys, columns = queries.
// ------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc5_dkv_kernel(const __grid_constant__ AttnMaps maps, const float* __restrict__ lse, const float* __restrict__ delta,
                    bf16* __restrict__ dqkv, int T, int H, float scale, float scale_log2) {
  using TL = Tile<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + TL::kBytes;
  const uint32_t sQ = sV + TL::kBytes;            // 2 tiles
  const uint32_t sDO = sQ + 2 * TL::kBytes;       // 2 tiles
  const uint32_t sPT = sDO + 2 * TL::kBytes;      // kPBytes: P^T_i   [keys x queries]
  const uint32_t sDST = sPT + kPBytes;            // kPBytes: dS^T_i
  float* sL = reinterpret_cast<float*>(smem + 6 * TL::kBytes + 2 * kPBytes);   // [256] lse (exp2 domain)
  float* sD = sL + 256;                                                          // [256] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
  uint64_t* bar_kv = bars;         // K, V landed
  uint64_t* bar_q = bars + 1;      // [2] Q_i, dO_i landed
  uint64_t* bar_s = bars + 3;      // [2] S^T_i, dP^T_i complete
  uint64_t* bar_p = bars + 5;      // [2] P^T_i, dS^T_i written
  uint64_t* bar_acc = bars + 7;    // [2] dV += P^T_i dO_i and dK += dS^T_i Q_i retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nblk = T / kRows;
  const int row0 = b * T + kb * kRows;

  if (threadIdx.x == 128) {
    mbar_init(bar_kv, 1);
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bar_q[j], 1);
      mbar_init(&bar_s[j], 1);
      mbar_init(&bar_p[j], 128);
      mbar_init(&bar_acc[j], 1);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<1>(tmem_slot, 512);
  // per-query statistics of this (batch, head)
  for (int i = threadIdx.x; i < T; i += kThreads) {
    sL[i] = lse[((int64_t)b * H + h) * T + i] * 1.4426950408889634f;
    sD[i] = delta[((int64_t)b * H + h) * T + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_kv, 2 * TL::kBytes);
      load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bar_kv, sK, H + h, row0);
      load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bar_kv, sV, 2 * H + h, row0);
      for (int i = 0; i < nblk; ++i) {
        mbar_expect_tx(&bar_q[i], 2 * TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &bar_q[i], sQ + i * TL::kBytes, h, b * T + i * kRows);
        load_tile<HD>(&maps.do_main, &maps.do_tail, &bar_q[i], sDO + i * TL::kBytes, h, b * T + i * kRows);
      }
      constexpr uint32_t idesc_s = make_idesc(128, 128, 0, 0);
      mbar_wait(bar_kv, 0);
      for (int i = 0; i < nblk; ++i) {
        mbar_wait(&bar_q[i], 0);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TL::kND / 16; ++ks)     // S^T_i = K Q_i^T
          umma(tmem + i * 256, desc_k<HD>(sK, ks), desc_k<HD>(sQ + i * TL::kBytes, ks), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < TL::kND / 16; ++ks)     // dP^T_i = V dO_i^T
          umma(tmem + i * 256 + 128, desc_k<HD>(sV, ks), desc_k<HD>(sDO + i * TL::kBytes, ks), idesc_s, ks > 0);
        commit(&bar_s[i]);
      }
      constexpr uint32_t idesc_o64 = make_idesc(128, 64, 0, 1);
      constexpr uint32_t idesc_o16 = make_idesc(128, 16, 0, 1);
      for (int i = 0; i < nblk; ++i) {                // dV -> columns 0.. (dead S^T_0), dK -> columns 128.. (dead dP^T_0)
        mbar_wait(&bar_p[i], 0);
        tc_fence_after();
        const uint32_t q = sQ + i * TL::kBytes, g = sDO + i * TL::kBytes;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma(tmem, desc_p(sPT, ks), desc_mn_main(g, ks), idesc_o64, i > 0 || ks > 0);
          if (TL::kTail) umma(tmem + 64, desc_p(sPT, ks), desc_mn_tail<HD>(g, ks), idesc_o16, i > 0 || ks > 0);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma(tmem + 128, desc_p(sDST, ks), desc_mn_main(q, ks), idesc_o64, i > 0 || ks > 0);
          if (TL::kTail) umma(tmem + 192, desc_p(sDST, ks), desc_mn_tail<HD>(q, ks), idesc_o16, i > 0 || ks > 0);
        }
        commit(&bar_acc[i]);
      }
    }
    __syncwarp();
  } else {
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int i = 0; i < nblk; ++i) {
      mbar_wait(&bar_s[i], 0);
      tc_fence_after();
      if (i > 0) mbar_wait(&bar_acc[i - 1], 0);   // the MMAs have finished reading P^T_{i-1} / dS^T_{i-1}
      const float* L = sL + i * kRows;
      const float* Dl = sD + i * kRows;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float s[32], dp[32];
        tmem_ld32_nowait(trow + i * 256 + c0, s);
        tmem_ld32_nowait(trow + i * 256 + 128 + c0, dp);
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
        store_p32(sPT, row, c0, s);
        store_p32(sDST, row, c0, dp);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar_p[i]);
    }
    mbar_wait(&bar_acc[nblk - 1], 0);
    tc_fence_after();
    const int tok = kb * kRows + row;
    bf16* outk = dqkv + ((int64_t)b * T + tok) * (3LL * H * HD) + (int64_t)H * HD + h * HD;
    store_acc_row<HD>(trow, 1.f, outk + (int64_t)H * HD);      // dV
    store_acc_row<HD>(trow + 128, scale, outk);                // dK
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<1>(tmem, 512);
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

// [rows, heads, hd] bf16, hd contiguous: box = 128 rows x 1 head x (64 cols, 128B swizzle | 16 cols, 32B swizzle)
int make_map3(CUtensorMap* map, const void* ptr, int64_t rows, int heads, int hd, bool tail) {
  EncodeTiledFn enc = encode_fn();
  REED_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)hd, (cuuint64_t)heads, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)hd * 2, (cuuint64_t)heads * hd * 2};
  cuuint32_t box[3] = {tail ? 16u : 64u, 1u, (cuuint32_t)kRows};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, tail ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REED_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d) rows=%lld heads=%d hd=%d", (int)r,
               (long long)rows, heads, hd);
  return 0;
}

template <int HD> constexpr int fwd_smem() { return 1024 + 5 * Tile<HD>::kBytes + 128; }
template <int HD> constexpr int dq_smem() { return 1024 + 6 * Tile<HD>::kBytes + kPBytes + 128; }
template <int HD> constexpr int dkv_smem() { return 1024 + 6 * Tile<HD>::kBytes + 2 * kPBytes + 2048 + 128; }

template <int HD>
int fwd_launch(const void* qkv, void* o, float* lse, int B, int T, int H, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_tc5_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem<HD>()));
    done = true;
  }
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (make_map3(&maps.qkv_main, qkv, (int64_t)B * T, 3 * H, HD, false)) return 1;
  if (Tile<HD>::kTail && make_map3(&maps.qkv_tail, qkv, (int64_t)B * T, 3 * H, HD, true)) return 1;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  attn_tc5_fwd_kernel<HD><<<dim3(T / kRows, H, B), kThreads, fwd_smem<HD>(), st>>>(maps, (bf16*)o, lse, T, H, scale_log2);
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
  if (make_map3(&maps.qkv_main, qkv, rows, 3 * H, HD, false)) return 1;
  if (make_map3(&maps.o_main, o, rows, H, HD, false)) return 1;
  if (make_map3(&maps.do_main, d_o, rows, H, HD, false)) return 1;
  if (Tile<HD>::kTail) {
    if (make_map3(&maps.qkv_tail, qkv, rows, 3 * H, HD, true)) return 1;
    if (make_map3(&maps.o_tail, o, rows, H, HD, true)) return 1;
    if (make_map3(&maps.do_tail, d_o, rows, H, HD, true)) return 1;
  }
  const float scale = 1.f / sqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(T / kRows, H, B);
  attn_tc5_dq_kernel<HD><<<grid, kThreads, dq_smem<HD>(), st>>>(maps, lse, (bf16*)dqkv, delta, T, H, scale, scale_log2);
  attn_tc5_dkv_kernel<HD><<<grid, kThreads, dkv_smem<HD>(), st>>>(maps, lse, delta, (bf16*)dqkv, T, H, scale, scale_log2);
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
