// Flash-style fused attention on tcgen05 / TMEM for any sequence of T = 128 n tokens (the 256-token and 1024-token patch
// sequences of the SiT hot path), head_dim 64 or 72; forward here, single-pass backward in attention_fa_bwd.cu.
//
// Reference semantics: timm Attention.forward with fused_attn (F.scaled_dot_product_attention, scale hd^-0.5), imported at
// /root/reference/image/models/sit.py:13 and called at sit.py:134.  Reads Q/K/V straight out of the packed qkv GEMM output
// [B,T,3,H,hd] through 3-D tensor maps, writes the context [B,T,H,hd] with TMA stores and the row log-sum-exp [B,H,T].
//
// Forward design (persistent, one CTA per SM, 320 threads):
//   * work item = TWO 128-row query tiles of one (batch, head); the key/value blocks of that head stream through a ring of
//     TMA stages ONCE for both tiles (the round-1 kernel fetched K and V once per query tile and prefetched nothing).
//   * warp 8 = TMA producer: Q double-buffered per item, K_j / V_j ring; loads of the next item run under the current one.
//   * warp 9 = MMA issuer (warp-uniform, one elected lane).  Per key block j and tile i:  S_i = Q_i K_j^T into TMEM, then
//     O_i += P_i V_j with P_i read FROM TENSOR MEMORY (tcgen05.mma A-operand in TMEM): the softmax warps overwrite S_i in
//     place with bf16 P_i through tcgen05.st, so probabilities never touch shared memory.  Issue order
//     PV_0(j) S_0(j+1) PV_1(j) S_1(j+1) keeps one tile's softmax running while the other tile's MMAs execute.
//   * warps 0-3 / 4-7 = softmax group of tile 0 / 1 (thread = query row = TMEM lane).  Online softmax in the exp2 domain
//     with a LAZY running maximum: the reference maximum of a row only moves when a block's maximum exceeds it by more
//     than 8 (P <= 2^8, exact in fp32 / bf16 range), so the O_i rescale (TMEM load-multiply-store) is rare; the final
//     normalisation uses the same reference, so the result is the exact softmax.  The row sum l = sum_k P is a sixteenth
//     ... an extra N = 16 MMA per k-step against a constant ones tile (O[:, hd'] += P . 1): 128 adds per row leave the
//     MUFU-bound softmax warps for the tensor pipe, and l is built from the same bf16 P the numerator uses.
//   * epilogue per tile: O_i / l -> bf16 -> swizzled staging tile -> TMA store; lse = (m + log2 l) ln 2.
//
// TMEM map (512 columns): S_0/P_0 0..127 | S_1/P_1 128..255 | O_0 256..335, row sum 336 | O_1 384..463, row sum 464.
#include <cuda.h>
#include "attention_fa.cuh"

namespace reed {
namespace {

using namespace fa;

constexpr int kFwdThreads = 320;
constexpr uint32_t kColS = 0, kColO = 256;          // + 128 per query tile
constexpr float kRescaleThreshold = 8.f;            // log2 units

template <int HD>
struct FwdCfg {
  using TL = Tile<HD>;
  static constexpr int kStages = HD > 64 ? 5 : 7;                     // K/V ring
  static constexpr int kOffQ = 0;                                     // [2 sets][2 tiles]
  static constexpr int kOffOut = 4 * TL::kBytes;                      // [2 tiles] output staging
  static constexpr int kOffKV = 6 * TL::kBytes;
  static constexpr int kOffOnes = kOffKV + kStages * TL::kBytes;      // [16 keys x 16] SWIZZLE_32B tile, column 0 = 1
  static constexpr int kOffBar = kOffOnes + 1024;
  // barriers: q_full[4] q_empty[4] kv_full[S] kv_empty[S] s_full[2] p_full[2] o_full[2]
  static constexpr int kNumBars = 8 + 2 * kStages + 6;
  static constexpr int kTotal = 1024 + kOffBar + kNumBars * 8 + 16;
  static_assert(kTotal <= 232448, "shared memory budget");
};

template <int HD>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fa_fwd_kernel(const __grid_constant__ AttnMaps maps, float* __restrict__ lse, int T, int H, int num_items,
                   float scale_log2) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  using TL = Tile<HD>;
  using CF = FwdCfg<HD>;
  constexpr int NST = CF::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::kOffBar);
  uint64_t* q_full = bars;                 // [set * 2 + tile]
  uint64_t* q_empty = bars + 4;
  uint64_t* kv_full = bars + 8;
  uint64_t* kv_empty = kv_full + NST;
  uint64_t* s_full = kv_empty + NST;       // [tile]
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = T / kRows;                       // key blocks
  const int nq = T / kRows;                        // query tiles (1, or an even number: attn_fa_supported)
  const int npair = (nq + 1) >> 1;
  const int nt = nq >= 2 ? 2 : 1;                  // query tiles per work item
  const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 256) {
    for (int k = 0; k < 8; ++k) mbar_init(bars + k, 1);
    for (int k = 0; k < 2 * NST; ++k) mbar_init(kv_full + k, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(p_full + i, 128);
      mbar_init(o_full + i, 1);
    }
    fence_barrier_init();
    // the ones tile of the row-sum MMA: every k-step (16 keys) reads the same 512 bytes; row r = 32 bytes, its first
    // 16-byte chunk sits at (r >> 2 & 1) * 16 under SWIZZLE_32B
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + CF::kOffOnes);
    for (int w = 0; w < 128; ++w) ones[w] = 0u;
    for (int r = 0; r < 16; ++r) ones[r * 8 + ((r >> 2) & 1) * 4] = 0x00003F80u;      // bf16 1.0 in element 0
    fence_proxy_async();
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

  auto decode = [&](int n, int& b, int& h, int& qp) {
    const int it = (int)blockIdx.x + n * (int)gridDim.x;
    qp = it % npair;
    h = (it / npair) % H;
    b = it / (npair * H);
  };

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int kvc = 0;                                  // K/V tiles issued so far (ring position)
      for (int n = 0; n < my_items; ++n) {
        int b, h, qp;
        decode(n, b, h, qp);
        const int set = n & 1;
        const uint32_t qph = (uint32_t)(n >> 1) & 1u;
        for (int i = 0; i < nt; ++i) {
          if (n >= 2) mbar_wait(&q_empty[set * 2 + i], qph ^ 1u);     // item n-2's score MMAs have read the buffer
          mbar_expect_tx(&q_full[set * 2 + i], TL::kBytes);
          load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &q_full[set * 2 + i], sbase + CF::kOffQ + (set * 2 + i) * TL::kBytes, h,
                        b * T + (2 * qp + i) * kRows);
        }
        for (int j = 0; j < nkb; ++j) {
#pragma unroll
          for (int kv = 0; kv < 2; ++kv, ++kvc) {
            const int st = kvc % NST;
            const uint32_t ph = (uint32_t)(kvc / NST) & 1u;
            if (kvc >= NST) mbar_wait(&kv_empty[st], ph ^ 1u);
            mbar_expect_tx(&kv_full[st], TL::kBytes);
            load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, &kv_full[st], sbase + CF::kOffKV + st * TL::kBytes,
                          (kv == 0 ? H : 2 * H) + h, b * T + j * kRows);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform; elected lane issues)
    const bool leader = elect_one();
    const int total = my_items * nkb;               // key blocks over all items of this CTA
    auto q_addr = [&](int n, int i) { return sbase + CF::kOffQ + ((n & 1) * 2 + i) * TL::kBytes; };
    auto kv_addr = [&](int c) { return sbase + CF::kOffKV + (c % NST) * TL::kBytes; };
    // S_i of block g (item n = g / nkb, key block j = g % nkb); K tile of block g is ring tile 2g, V tile 2g + 1
    auto issue_scores = [&](int g, int i) {
      const int n = g / nkb, j = g % nkb;
      const bool last_tile = i == nt - 1;
      if (i == 0) {
        mbar_wait(&kv_full[(2 * g) % NST], (uint32_t)((2 * g) / NST) & 1u);
      }
      if (j == 0) mbar_wait(&q_full[(n & 1) * 2 + i], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      mma_scores<HD>(leader, tmem + kColS + i * 128, q_addr(n, i), kv_addr(2 * g));
      commit_if(leader, &s_full[i]);
      if (j == nkb - 1) commit_if(leader, &q_empty[(n & 1) * 2 + i]);     // last read of Q_i of this item
      if (last_tile) commit_if(leader, &kv_empty[(2 * g) % NST]);
    };
    if (total > 0)
      for (int i = 0; i < nt; ++i) issue_scores(0, i);
    for (int g = 0; g < total; ++g) {
      const int j = g % nkb;
      const uint32_t pph = (uint32_t)g & 1u;           // every tile sees every block: one p_full phase per block
      const int vst = (2 * g + 1) % NST;
      for (int i = 0; i < nt; ++i) {
        if (i == 0) mbar_wait(&kv_full[vst], (uint32_t)((2 * g + 1) / NST) & 1u);
        mbar_wait(&p_full[i], pph);
        tc_fence_after();
        mma_pv_ts<HD>(leader, tmem + kColO + i * 128, tmem + kColS + i * 128, kv_addr(2 * g + 1), j > 0, sbase + CF::kOffOnes);
        if (j == nkb - 1) commit_if(leader, &o_full[i]);
        if (i == nt - 1) commit_if(leader, &kv_empty[vst]);
        if (g + 1 < total) issue_scores(g + 1, i);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- softmax groups: thread = query row = TMEM lane
    const int i = warp >> 2;                        // query tile of the pair
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + kColS + i * 128;
    const uint32_t tO = tmem + lane_base + kColO + i * 128;
    const uint32_t sOut = sbase + CF::kOffOut + i * TL::kBytes;
    int sc = 0;                                     // score blocks of this tile consumed so far (phase of s_full / p_full)
    int oc = 0;                                     // items of this tile finished (phase of o_full)
    for (int n = 0; n < my_items; ++n) {
      int b, h, qp;
      decode(n, b, h, qp);
      if (i >= nt) break;                           // T = 128: one query tile, the second group has no work
      float m_used = 0.f;
      for (int j = 0; j < nkb; ++j, ++sc) {
        mbar_wait(&s_full[i], (uint32_t)sc & 1u);
        tc_fence_after();
        // the whole score row of the block -> registers (second pair of loads flies under the first half's maximum)
        float v[128];
        tmem_ld32_nowait(tS, v);
        tmem_ld32_nowait(tS + 32, v + 32);
        tmem_wait_ld();
        tmem_ld32_nowait(tS + 64, v + 64);
        tmem_ld32_nowait(tS + 96, v + 96);
        float mx = fmaxf(v[0], v[1]);
#pragma unroll
        for (int e = 2; e < 64; e += 2) mx = fmaxf(mx, fmaxf(v[e], v[e + 1]));
        tmem_wait_ld();
#pragma unroll
        for (int e = 64; e < 128; e += 2) mx = fmaxf(mx, fmaxf(v[e], v[e + 1]));
        const float m_blk = mx * scale_log2;
        if (j == 0) {
          m_used = m_blk;
        } else {
          const bool grow = m_blk > m_used + kRescaleThreshold;
          if (__any_sync(0xffffffffu, grow)) {
            // rare: move this row's reference maximum and rescale its accumulator row (and its running sum, column
            // kND) in TMEM.  PV of block j-1 has retired: s_full of block j was committed behind it.
            const float m_new = grow ? m_blk : m_used;
            const float alpha = ex2(m_used - m_new);
            m_used = m_new;
#pragma unroll
            for (int c0 = 0; c0 < TL::kND + 16; c0 += 16) {
              float x[16];
              tmem_ld16_nowait(tO + c0, x);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) x[e] *= alpha;
              tmem_st16(tO + c0, reinterpret_cast<const uint32_t*>(x));
            }
          }
        }
        // P = 2^(s * scale - m) -> bf16 pairs over the S row.  The row sum is NOT taken here: the PV MMA carries a ones
        // column (O[:, kND] += P . 1), so the 128 adds per row leave the softmax warps for the tensor pipe.
        const float neg_m = -m_used;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[c * 32 + e] = fmaf(v[c * 32 + e], scale_log2, neg_m);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[c * 32 + e] = ex2(v[c * 32 + e]);
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) pk[e >> 1] = pack2(v[c * 32 + e], v[c * 32 + e + 1]);
          tmem_st16(tS + c * 16, pk);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[i]);
      }
      // ---- epilogue of this tile: O / l -> bf16 -> staging -> TMA store
      mbar_wait(&o_full[i], (uint32_t)oc & 1u);
      ++oc;
      tc_fence_after();
      float l;
      {
        float x[16];
        tmem_ld16_nowait(tO + TL::kND, x);          // column kND = sum_k P (the ones column of the PV MMA)
        tmem_wait_ld();
        l = x[0];
      }
      const float inv = 1.f / l;
      const int tok = (2 * qp + i) * kRows + row;
      lse[((int64_t)b * H + h) * T + tok] = (m_used + log2f(l)) * 0.6931471805599453f;
      if ((warp & 3) == 0 && lane == 0) tma_store_wait_read();      // the previous store of this group has read sOut
      asm volatile("bar.sync %0, 128;" ::"r"(1 + i) : "memory");
#pragma unroll
      for (int c0 = 0; c0 < TL::kND; c0 += 16) {
        float x[16];
        tmem_ld16_nowait(tO + c0, x);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) x[e] *= inv;
        stage_out8<HD>(sOut, row, c0 >> 3, x);
        if (c0 + 8 < HD) stage_out8<HD>(sOut, row, (c0 >> 3) + 1, x + 8);
      }
      tc_fence_before();                             // O_i is read out: the next item's first PV_i may overwrite it
      fence_proxy_async();                           // (ordered before this thread's next p_full arrive)
      asm volatile("bar.sync %0, 128;" ::"r"(1 + i) : "memory");
      if ((warp & 3) == 0 && lane == 0) {
        store_tile<HD>(&maps.out_main, &maps.out_tail8, sOut, h, b * T + (2 * qp + i) * kRows);
        tma_store_commit();
      }
    }
    if ((warp & 3) == 0 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem, 512);
}

template <int HD>
int fwd_launch(const void* qkv, void* o, float* lse, int B, int T, int H, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(attn_fa_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<HD>::kTotal));
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
  const int nq = T / kRows;
  const int items = ((nq + 1) / 2) * H * B;
  const int grid = items < sm_count() ? items : sm_count();
  attn_fa_fwd_kernel<HD><<<grid, kFwdThreads, FwdCfg<HD>::kTotal, st>>>(maps, lse, T, H, items, scale_log2);
  REED_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// T = 128 (one query tile per item) or a multiple of 256 (query tiles in pairs)
bool attn_fa_supported(int T, int hd) { return (T == 128 || (T >= 256 && T % 256 == 0)) && (hd == 64 || hd == 72); }

int attn_fa_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int hd, cudaStream_t st) {
  if (B == 0) return 0;
  if (hd == 64) return fwd_launch<64>(qkv, o, lse, B, T, H, st);
  if (hd == 72) return fwd_launch<72>(qkv, o, lse, B, T, H, st);
  return fail("tcgen05 attention: head_dim %d unsupported", hd);
}

}  // namespace reed
