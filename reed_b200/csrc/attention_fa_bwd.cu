// Single-pass attention backward on tcgen05 / TMEM: dQ, dK and dV of one (batch, head) in ONE kernel, the probabilities
// recomputed once.  Autograd of timm Attention.forward with fused_attn (/root/reference/image/models/sit.py:13,134).
//
// Decomposition.  A thread-block CLUSTER of n = T / 128 CTAs owns one (batch, head); CTA rank r owns key block r for the
// whole item: K_r and V_r stay in shared memory, dK_r and dV_r accumulate in tensor memory over the n query blocks.  At step
// s the CTA works on query block i = (r + 1 + s) mod n, so that at any time the n CTAs hold n different query blocks:
//     S^T  = K_r Q_i^T                (lanes = keys, columns = queries)          tcgen05.mma, operands by TMA
//     dP^T = V_r dO_i^T
//     P^T  = 2^(S^T scale - lse_i),  dS^T = P^T (dP^T - delta_i)                 two elementwise groups, thread = key row
//     dV_r += P^T dO_i               P^T read from TENSOR MEMORY (written in place over S^T with tcgen05.st)
//     dK_r += dS^T Q_i               dS^T read from TENSOR MEMORY too (written over dP^T): an A operand in shared memory
//                                    would be re-read for the 16-wide head_dim tail MMA and make these MMAs smem-bound
//     dQ_i(partial) = dS K_r         dS^T also goes to shared memory once, read as an MN-major A operand (no transposed copy)
// Issue order per step: dV, S^T(next), dK, dP^T(next), dQ - the next step's exponentials (which need S^T only) start while
// dK / dP^T / dQ execute, and P^T / dS^T are handed over separately so dV does not wait for the dS arithmetic.
// The partial dQ_i of the n key blocks is summed around a RING over distributed shared memory: the CTA adds its partial to
// the running sum it received from rank r+1 and forwards it to rank r-1, which works on query block i one step later; at
// its last step (i = r) a CTA holds the complete dQ_r and stores it.  No fp32 dQ workspace, no atomics, no second pass:
// HBM traffic is qkv + dO in, dqkv out.  delta = rowsum(dO * O) comes from a small pre-kernel ([B,H,T] fp32).
//
// Persistent clusters: each cluster loops over (batch, head) items; K/V of the next item and Q/dO of the next step are
// prefetched by TMA into the other half of two double buffers while the current step computes.
//
// Warp roles (512 threads): warps 0-3 / 4-7 elementwise groups (queries 0..63 / 64..127 of the tile), warps 8-11 drain group
// (dQ ring + dK/dV epilogue, thread = TMEM lane), warp 12 Q/dO producer, warp 13 MMA issuer, warp 14 K/V + statistics
// producer, warp 15 store thread (TMA stores of the staged results and the waits behind them).  One warp per stream: lanes
// of ONE warp spinning on different mbarriers starve each other for thousands of cycles.
// TMEM map: S^T/P^T 0..127 | dP^T 128..255 | dV | dK | dQ (kND columns each) from 256.
#include <cuda.h>
#ifdef REED_ATTN_DEBUG
#define REED_MBAR_SPINS (1ull << 20)
#endif
#include "attention_fa.cuh"

namespace reed {

// clock64 trace of CTA 0 (debug builds: make EXTRA=-DREED_ATTN_TRACE; read with profiles/attn_trace_bwd.py)
#ifdef REED_ATTN_TRACE
__device__ unsigned long long g_trace[5][2048];
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

using namespace fa;

constexpr int kBwdThreads = 512;
constexpr int kMaxT = 1024;

template <int HD>
struct BwdCfg {
  using TL = Tile<HD>;
  static constexpr int kOffKV = 0;                                   // [2 sets][K, V]
  static constexpr int kOffQdO = 4 * TL::kBytes;                     // [2 bufs][Q, dO]
  static constexpr int kOffDS = 8 * TL::kBytes;                      // dS^T: two [128 x 64] K-major SWIZZLE_128B tiles
  static constexpr int kOffRecv = kOffDS + kPBytes;                  // incoming dQ running sum, bf16, TMA-store staging layout
  static constexpr int kOffStat = kOffRecv + TL::kBytes;             // lse (log2 domain) and delta of the item(s): 8 KB
  static constexpr int kStatFloats = 2048;
  static constexpr int kOffBar = kOffStat + kStatFloats * 4;
  static constexpr int kNumBars = 26;
  static constexpr int kTotal = 1024 + kOffBar + kNumBars * 8 + 16;
  static constexpr uint32_t kColST = 0, kColDP = 128, kColDV = 256, kColDK = 256 + TL::kND, kColDQ = 256 + 2 * TL::kND;
  static_assert(kTotal <= 232448, "shared memory budget");
  static_assert(256 + 3 * TL::kND <= 512, "tensor memory budget");
};

// barrier indices
enum {
  kKvFull = 0,      // [2] K_r, V_r of item parity landed (TMA)
  kKvEmpty = 2,     // [2] dK/dV stores have read the staging that aliases the set (drain thread)
  kQdoFull = 4,     // [2] Q_i, dO_i landed
  kQdoEmpty = 6,    // [2] every MMA of the step has read them (commit)
  kSFull = 8,       // S^T complete (commit)
  kDpFull = 9,      // dP^T complete (commit)
  kPFull = 10,      // P^T written to tensor memory over S^T (256 elementwise threads)
  kDstFull = 11,    // dS^T written to tensor memory over dP^T (256)
  kDsFull = 12,     // dS^T written to shared memory (256)
  kDsFree = 13,     // the dQ MMA of the step has read the shared-memory dS^T (commit)
  kDqFull = 14,     // partial dQ complete (commit)
  kDqDrained = 15,  // drain group has read it out of TMEM (128)
  kDkvFull = 16,    // dK, dV of the item complete (commit)
  kDvDrained = 17,  // drain group has read dV out of tensor memory (128): the next item's first dV MMA may overwrite it
  kDkDrained = 18,  // ... and dK (128)
  kStatFull = 19,   // [2] lse / delta of an item are in shared memory (bulk copies: 1 arrival + bytes)
  kRecvFull = 21,   // running dQ sum from rank r+1 arrived (st.async from the neighbour: 1 local arrival + bytes)
  kSendCredit = 22, // rank r-1 has consumed our previous message (1 remote arrival)
  kStageFull = 23,  // dQ / dK / dV of the item are staged for their TMA stores (128)
  kStoreDone = 24,  // ... and the stores have read the staging (store thread)
};

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  uint64_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > REED_MBAR_SPINS) {
      printf("reed attention bwd: cluster mbarrier wait timed out (block %d thread %d, barrier slot %u, parity %u)\n", blockIdx.x,
             threadIdx.x, (addr & 1023u) >> 3, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// 8 values -> bf16 -> 16 bytes of a peer CTA's shared memory, asynchronously: the store reports its bytes to the mbarrier
// `bar_cluster` of the same peer when it has landed, so the sender needs no fence and no arrive of its own
__device__ __forceinline__ void st_async128_cluster(uint32_t addr_cluster, const float* v, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr_cluster), "r"(pack2(v[0], v[1])), "r"(pack2(v[2], v[3])), "r"(pack2(v[4], v[5])),
                 "r"(pack2(v[6], v[7])), "r"(bar_cluster)
               : "memory");
}
// 1-D bulk copy global -> shared memory of `bytes` (multiple of 16), completing on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// byte offset of 8-column chunk `chunk` of row `row` inside a TMA-store staging tile (stage_out8's layout)
template <int HD>
__device__ __forceinline__ uint32_t stage_offset(int row, int chunk) {
  return chunk < 8 ? (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)) : (uint32_t)(Tile<HD>::kMain + row * 16);
}
__device__ __forceinline__ void add_bf16x8(float* x, uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    x[2 * e] += __low2float(p[e]);
    x[2 * e + 1] += __high2float(p[e]);
  }
}

// D[128 x hd] (+)= A . Z with A = dS^T tiles in shared memory read K-major (rows = M) or MN-major (rows = K):
//   kMnA = false: D[key, d]   += sum_q dS^T[key, q] Z[q, d]      (dK:  Z = Q_i)
//   kMnA = true : D[query, d] += sum_k dS^T[k, query] Z[k, d]    (dQ:  Z = K_r)
// Z is a TMA-staged [128 x hd] tile read MN-major; N = 64 main + 16 tail.
template <int HD, bool kMnA>
__device__ __forceinline__ void mma_ds(bool leader, uint32_t tmem_d, uint32_t ds, uint32_t tile_z, bool accumulate) {
  constexpr uint32_t idesc64 = make_idesc(128, 64, kMnA ? 1 : 0, 1);
  constexpr uint32_t idesc16 = make_idesc(128, 16, kMnA ? 1 : 0, 1);
  const uint32_t lz = desc_lo(tile_z);
  const uint32_t la = kMnA ? desc_lo_lbo(ds, kRows * 128) : desc_lo(ds);
  if (leader) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      // K-major: k-step = 16 queries = 32 bytes inside the 128-byte row of tile ks >> 2; MN-major: 16 key rows = 2048 bytes
      const uint32_t a_lo = kMnA ? la + ks * (2048 >> 4) : la + (ks >> 2) * (kRows * 128 >> 4) + (ks & 3) * 2;
      const uint64_t da = mk_desc(a_lo, kHiSw128);
      const uint32_t acc = (accumulate || ks > 0) ? 1u : 0u;
      umma_bf16<1>(tmem_d, da, mk_desc(lz + ks * (2048 >> 4), kHiSw128), idesc64, acc);
      if (Tile<HD>::kTail)
        umma_bf16<1>(tmem_d + 64, da, mk_desc(lz + (Tile<HD>::kMain >> 4) + ks * (512 >> 4), kHiSw32), idesc16, acc);
    }
  }
}
// dV[key, d] (+)= sum_q P^T[key, q] dO[q, d], P^T in tensor memory: queries 0..63 at columns 0..31 of `tmem_p`, queries
// 64..127 at columns 64..95 (each elementwise group packs its half over the S^T columns it has consumed)
template <int HD>
__device__ __forceinline__ void mma_dv_ts(bool leader, uint32_t tmem_d, uint32_t tmem_p, uint32_t tile_z, bool accumulate) {
  constexpr uint32_t idesc64 = make_idesc(128, 64, 0, 1);
  constexpr uint32_t idesc16 = make_idesc(128, 16, 0, 1);
  const uint32_t lz = desc_lo(tile_z);
  if (leader) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t pa = tmem_p + (ks >> 2) * 64 + (ks & 3) * 8;
      const uint32_t acc = (accumulate || ks > 0) ? 1u : 0u;
      umma_ts(tmem_d, pa, mk_desc(lz + ks * (2048 >> 4), kHiSw128), idesc64, acc);
      if (Tile<HD>::kTail)
        umma_ts(tmem_d + 64, pa, mk_desc(lz + (Tile<HD>::kMain >> 4) + ks * (512 >> 4), kHiSw32), idesc16, acc);
    }
  }
}

// 32 packed bf16 pairs-of-columns (16 words: columns c0 .. c0 + 31 of row `row`) -> the dS^T tiles (store_p32's layout)
__device__ __forceinline__ void store_p32_packed(uint32_t base, int row, int c0, const uint32_t* pk) {
  const uint32_t tile = base + (c0 >> 6) * (kRows * 128) + row * 128;
  const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + (((chunk0 + q) ^ (row & 7)) << 4)), "r"(pk[4 * q]),
                 "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                 : "memory");
}

template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_fa_bwd_kernel(const __grid_constant__ AttnMaps maps, const float* __restrict__ lse, const float* __restrict__ delta,
                   int T, int H, int num_items, float scale, float scale_log2) {
  pdl_launch();
  using TL = Tile<HD>;
  using CF = BwdCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + CF::kNumBars);
  float* stat = reinterpret_cast<float*>(smem + CF::kOffStat);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nst = T / kRows;                          // steps per item = cluster size = key blocks = query blocks
  const uint32_t rank = nst > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = (int)blockIdx.x / nst, num_clusters = (int)gridDim.x / nst;
  const int my_items = (num_items - cluster_id + num_clusters - 1) / num_clusters;
  const int stat_bufs = (2 * 2 * T <= CF::kStatFloats) ? 2 : 1;      // lse + delta of two items fit (T <= 512)

  if (threadIdx.x == 0) {
    for (int k = 0; k < CF::kNumBars; ++k) {
      const bool by_ew = k == kPFull || k == kDstFull || k == kDsFull;
      const bool by_drain = k == kDqDrained || k == kDvDrained || k == kDkDrained || k == kStageFull;
      mbar_init(bars + k, by_ew ? 256u : (by_drain ? 128u : 1u));
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.qkv_main);
    tma_prefetch_desc(&maps.do_main);
    tma_prefetch_desc(&maps.out_main);
    if (TL::kTail) {
      tma_prefetch_desc(&maps.qkv_tail);
      tma_prefetch_desc(&maps.do_tail);
      tma_prefetch_desc(&maps.out_tail8);
    }
  }
  if (warp == 13) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (nst > 1) cluster_sync_all();                    // every CTA's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto item_bh = [&](int n, int& b, int& h) {
    const int it = cluster_id + n * num_clusters;
    h = it % H;
    b = it / H;
  };
  auto q_block = [&](int s) { return (int)((rank + 1 + (uint32_t)s) % (uint32_t)nst); };
  auto kv_slot = [&](int n, int which) { return sbase + CF::kOffKV + ((n & 1) * 2 + which) * TL::kBytes; };
  auto qdo_slot = [&](int g, int which) { return sbase + CF::kOffQdO + ((g & 1) * 2 + which) * TL::kBytes; };

  if (warp == 12 || warp == 14 || warp == 15) {
    // ------------------------------------------------------------------ TMA streams, one warp each (lane 0 works)
    // (independent streams: the K/V of the next item must not queue behind a Q/dO slot that frees a step later)
    if (warp == 12 && lane == 0) {
      TRACE_DECL(4, true);
      int g = 0;
      for (int n = 0; n < my_items; ++n) {
        int b, h;
        item_bh(n, b, h);
        for (int s = 0; s < nst; ++s, ++g) {
          const int buf = g & 1;
          TRACE(1);
          if (g >= 2) mbar_wait(bars + kQdoEmpty + buf, ((uint32_t)(g >> 1) & 1u) ^ 1u);
          TRACE(2);
          const int row0 = b * T + q_block(s) * kRows;
          mbar_expect_tx(bars + kQdoFull + buf, 2 * TL::kBytes);
          load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bars + kQdoFull + buf, qdo_slot(g, 0), h, row0);
          load_tile<HD>(&maps.do_main, &maps.do_tail, bars + kQdoFull + buf, qdo_slot(g, 1), h, row0);
        }
      }
    } else if (warp == 14 && lane == 0) {
      for (int n = 0; n < my_items; ++n) {
        int b, h;
        item_bh(n, b, h);
        const int set = n & 1;
        if (n >= 2) mbar_wait(bars + kKvEmpty + set, ((uint32_t)(n >> 1) & 1u) ^ 1u);
        mbar_expect_tx(bars + kKvFull + set, 2 * TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bars + kKvFull + set, kv_slot(n, 0), H + h, b * T + (int)rank * kRows);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bars + kKvFull + set, kv_slot(n, 1), 2 * H + h, b * T + (int)rank * kRows);
        // lse and delta rows of the item.  Two buffers (T <= 512): free once item n-2 is done, which the K/V wait above
        // implies; one buffer (T = 1024): free once the previous item's last MMAs - hence all its elementwise work - are done
        if (stat_bufs == 1 && n >= 1) mbar_wait(bars + kDkvFull, (uint32_t)(n - 1) & 1u);
        const int sb = n % stat_bufs;
        const int64_t base = ((int64_t)b * H + h) * T;
        mbar_expect_tx(bars + kStatFull + sb, 2 * T * 4);
        bulk_load(smem_u32(stat + sb * 2 * T), lse + base, T * 4, bars + kStatFull + sb);
        bulk_load(smem_u32(stat + sb * 2 * T + T), delta + base, T * 4, bars + kStatFull + sb);
      }
    } else if (warp == 15 && lane == 0) {
      // store thread: TMA stores of the staged dQ / dK / dV and everything that has to wait for them, off the drain group
      const uint32_t right = (rank + 1u) % (uint32_t)nst;
      const uint32_t remote_credit = nst > 1 ? mapa_u32(smem_u32(bars + kSendCredit), right) : 0u;
      for (int n = 0; n < my_items; ++n) {
        int b, h;
        item_bh(n, b, h);
        mbar_wait(bars + kStageFull, (uint32_t)n & 1u);
        const int row0 = b * T + (int)rank * kRows;
        // dK / dV first, as their own bulk group: the K/V set they are staged in is what the next-but-one item's operand
        // prefetch waits for; the dQ store (receive buffer) follows
        store_tile<HD>(&maps.out_main, &maps.out_tail8, kv_slot(n, 1), 2 * H + h, row0);
        store_tile<HD>(&maps.out_main, &maps.out_tail8, kv_slot(n, 0), H + h, row0);
        tma_store_commit();
        store_tile<HD>(&maps.out_main, &maps.out_tail8, sbase + CF::kOffRecv, h, row0);
        tma_store_commit();
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the dK / dV stores have read their staging
        mbar_arrive(bars + kKvEmpty + (n & 1));
        tma_store_wait_read();                        // ... and the dQ store
        if (nst > 1) mbar_arrive_remote_release(remote_credit);       // the right neighbour may send into our buffer again
        mbar_arrive(bars + kStoreDone);
      }
      tma_store_wait_all();
    }
    __syncwarp();
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform; elected lane issues)
    const bool leader = elect_one();
    TRACE_DECL(2, lane == 0);
    const int G = my_items * nst;
    const uint32_t sDS = sbase + CF::kOffDS;
    auto wait_operands = [&](int g) {
      const int n = g / nst, s = g % nst;
      if (s == 0) mbar_wait(bars + kKvFull + (n & 1), (uint32_t)(n >> 1) & 1u);
      mbar_wait(bars + kQdoFull + (g & 1), (uint32_t)(g >> 1) & 1u);
      tc_fence_after();
    };
    auto issue_st = [&](int g) {       // S^T = K_r Q_i^T
      mma_scores<HD>(leader, tmem + CF::kColST, kv_slot(g / nst, 0), qdo_slot(g, 0));
      commit_if(leader, bars + kSFull);
    };
    auto issue_dp = [&](int g) {       // dP^T = V_r dO_i^T
      mma_scores<HD>(leader, tmem + CF::kColDP, kv_slot(g / nst, 1), qdo_slot(g, 1));
      commit_if(leader, bars + kDpFull);
    };
    if (G > 0) {
      wait_operands(0);
      issue_st(0);
      issue_dp(0);
    }
    for (int g = 0; g < G; ++g) {
      const int n = g / nst, s = g % nst;
      mbar_wait(bars + kPFull, (uint32_t)g & 1u);
      TRACE(1);
      if (s == 0 && n > 0) mbar_wait(bars + kDvDrained, (uint32_t)(n - 1) & 1u);    // previous item's dV has been read out
      tc_fence_after();
      mma_dv_ts<HD>(leader, tmem + CF::kColDV, tmem + CF::kColST, qdo_slot(g, 1), s > 0);       // dV += P^T dO_i
      if (g + 1 < G) {                   // behind dV the S^T / P^T columns are free again (in-order pipe)
        TRACE(6);
        wait_operands(g + 1);
        TRACE(7);
        issue_st(g + 1);
      }
      mbar_wait(bars + kDstFull, (uint32_t)g & 1u);
      TRACE(2);
      if (s == 0 && n > 0) mbar_wait(bars + kDkDrained, (uint32_t)(n - 1) & 1u);    // ... and its dK
      tc_fence_after();
      mma_dv_ts<HD>(leader, tmem + CF::kColDK, tmem + CF::kColDP, qdo_slot(g, 0), s > 0);       // dK += dS^T Q_i
      if (g + 1 < G) issue_dp(g + 1);    // behind dK the dP^T / dS^T columns are free again
      mbar_wait(bars + kDsFull, (uint32_t)g & 1u);
      TRACE(3);
      if (g > 0) mbar_wait(bars + kDqDrained, (uint32_t)(g - 1) & 1u);
      tc_fence_after();
      TRACE(4);
      mma_ds<HD, true>(leader, tmem + CF::kColDQ, sDS, kv_slot(n, 0), false);                   // dQ_i(partial) = dS K_r
      TRACE(5);
      commit_if(leader, bars + kDqFull);
      commit_if(leader, bars + kDsFree);
      commit_if(leader, bars + kQdoEmpty + (g & 1));
      if (s == nst - 1) commit_if(leader, bars + kDkvFull);
    }
    __syncwarp();
  } else if (warp < 8) {
    // ------------------------------------------------------------------ elementwise groups: thread = key row (TMEM lane)
    const int grp = warp >> 2;                        // queries [64 grp, 64 grp + 64) of the tile
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tST = tmem + lane_base + CF::kColST + grp * 64;
    const uint32_t tDP = tmem + lane_base + CF::kColDP + grp * 64;
    const uint32_t sDS = sbase + CF::kOffDS;
    TRACE_DECL(grp, (warp & 3) == 0 && lane == 0);
    int g = 0;
    for (int n = 0; n < my_items; ++n) {
      const int sb = n % stat_bufs;
      TRACE(6);
      mbar_wait(bars + kStatFull + sb, (uint32_t)(n / stat_bufs) & 1u);
      TRACE(7);
      const float* L = stat + sb * 2 * T;             // lse (natural log, as the forward wrote it)
      const float* Dl = L + T;
      for (int s = 0; s < nst; ++s, ++g) {
        const int q0 = q_block(s) * kRows + grp * 64;
        // phase 1: probabilities (needs S^T only).  P^T goes over the S^T columns this thread has consumed.
        TRACE(1);
        mbar_wait(bars + kSFull, (uint32_t)g & 1u);
        TRACE(2);
        tc_fence_after();
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float sv[32];
          tmem_ld32_nowait(tST + c * 32, sv);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            float4 l4 = *reinterpret_cast<const float4*>(L + q0 + c * 32 + q);           // same address in every lane: broadcast
            constexpr float kNegLog2e = -1.4426950408889634f;
            l4.x *= kNegLog2e, l4.y *= kNegLog2e, l4.z *= kNegLog2e, l4.w *= kNegLog2e;
            pk[c * 16 + (q >> 1)] = pack2(ex2(fmaf(sv[q], scale_log2, l4.x)), ex2(fmaf(sv[q + 1], scale_log2, l4.y)));
            pk[c * 16 + (q >> 1) + 1] = pack2(ex2(fmaf(sv[q + 2], scale_log2, l4.z)), ex2(fmaf(sv[q + 3], scale_log2, l4.w)));
          }
        }
        tmem_st16(tST, pk);
        tmem_st16(tST + 16, pk + 16);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bars + kPFull);                   // dV and the next step's S^T may go
        TRACE(3);
        // phase 2: dS^T = P^T (dP^T - delta), with the bf16 P the dV MMA sees; over the dP^T columns and to shared memory
        mbar_wait(bars + kDpFull, (uint32_t)g & 1u);
        tc_fence_after();
        uint32_t dsp[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float dp[32];
          tmem_ld32_nowait(tDP + c * 32, dp);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(Dl + q0 + c * 32 + q);
            const __nv_bfloat162 p01 = *reinterpret_cast<const __nv_bfloat162*>(&pk[c * 16 + (q >> 1)]);
            const __nv_bfloat162 p23 = *reinterpret_cast<const __nv_bfloat162*>(&pk[c * 16 + (q >> 1) + 1]);
            // dS without the softmax scale (applied to dQ / dK at the end)
            dsp[c * 16 + (q >> 1)] = pack2(__low2float(p01) * (dp[q] - d4.x), __high2float(p01) * (dp[q + 1] - d4.y));
            dsp[c * 16 + (q >> 1) + 1] = pack2(__low2float(p23) * (dp[q + 2] - d4.z), __high2float(p23) * (dp[q + 3] - d4.w));
          }
        }
        tmem_st16(tDP, dsp);
        tmem_st16(tDP + 16, dsp + 16);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bars + kDstFull);                 // dK and the next step's dP^T may go
        if (g > 0) mbar_wait(bars + kDsFree, (uint32_t)(g - 1) & 1u);    // dQ of the previous step has read the smem tiles
        TRACE(4);
        store_p32_packed(sDS, row, grp * 64, dsp);
        store_p32_packed(sDS, row, grp * 64 + 32, dsp + 16);
        fence_proxy_async();
        mbar_arrive(bars + kDsFull);
        TRACE(5);
      }
    }
  } else {
    // ------------------------------------------------------------------ drain group: thread = TMEM lane
    const int dw = warp - 8;
    const int row = dw * 32 + lane;
    const int tid = row;
    const uint32_t lane_base = (uint32_t)(dw * 32) << 16;
    const uint32_t sRecv = sbase + CF::kOffRecv;
    const uint32_t left = (rank + (uint32_t)nst - 1u) % (uint32_t)nst;      // we send to it
    const uint32_t right = (rank + 1u) % (uint32_t)nst;                    // we receive from it
    const uint32_t remote_recv = nst > 1 ? mapa_u32(sRecv, left) : 0u;
    const uint32_t remote_recv_full = nst > 1 ? mapa_u32(smem_u32(bars + kRecvFull), left) : 0u;
    const uint32_t remote_credit = nst > 1 ? mapa_u32(smem_u32(bars + kSendCredit), right) : 0u;
    TRACE_DECL(3, tid == 0);
    constexpr uint32_t kMsgBytes = kRows * HD * 2;     // one running dQ sum: 128 rows of head_dim bf16
    int g = 0, sends = 0, recvs = 0;
    for (int n = 0; n < my_items; ++n) {
      int b, h;
      item_bh(n, b, h);
      TRACE(13);
      for (int s = 0; s < nst; ++s, ++g) {
        const bool last = s == nst - 1;
        TRACE(1);
        mbar_wait(bars + kDqFull, (uint32_t)g & 1u);
        TRACE(2);
        tc_fence_after();
        constexpr int kChunks = (HD + 15) / 16;
        const uint32_t sK = kv_slot(n, 0), sV = kv_slot(n, 1);        // dead operands after the last step: dK / dV staging
        if (last && n > 0) mbar_wait(bars + kStoreDone, (uint32_t)(n - 1) & 1u);     // the staging buffers are free again
        if (last) {
          // dV first: the next item's first dV MMA waits for these columns (everything else of the epilogue is off the
          // critical path).  kDkvFull and kDqFull are committed together, so no extra wait is paid here.
          mbar_wait(bars + kDkvFull, (uint32_t)n & 1u);
          tc_fence_after();
          float v[kChunks][16];
#pragma unroll
          for (int c = 0; c < kChunks; ++c) tmem_ld16_nowait(tmem + lane_base + CF::kColDV + c * 16, v[c]);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(bars + kDvDrained);
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            stage_out8<HD>(sV, row, 2 * c, v[c]);
            if (c * 16 + 8 < HD) stage_out8<HD>(sV, row, 2 * c + 1, v[c] + 8);
          }
        }
        if (s > 0) {
          if (tid == 0) mbar_expect_tx(bars + kRecvFull, kMsgBytes);      // the neighbour's st.async stores carry the bytes
          mbar_wait_cluster(bars + kRecvFull, (uint32_t)recvs & 1u);
          ++recvs;
        }
        TRACE(3);
        // consume first (partial dQ out of tensor memory + the running sum that arrived), release both, and only then
        // wait for room at the left neighbour: waiting for the send credit before consuming deadlocks a ring of n > 2
        float x[kChunks][16];
#pragma unroll
        for (int c = 0; c < kChunks; ++c) tmem_ld16_nowait(tmem + lane_base + CF::kColDQ + c * 16, x[c]);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(bars + kDqDrained);
        if (s > 0) {
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            add_bf16x8(x[c], sRecv + stage_offset<HD>(row, 2 * c));
            if (c * 16 + 8 < HD) add_bf16x8(x[c] + 8, sRecv + stage_offset<HD>(row, 2 * c + 1));   // hd 72: 8 real columns in the last chunk
          }
        }
        TRACE(4);
        if (!last) {
          if (s > 0) {                                // our receive buffer has been read by all 128 threads: credit the sender
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) mbar_arrive_remote_release(remote_credit);
          }
          if (sends > 0) mbar_wait_cluster(bars + kSendCredit, (uint32_t)(sends - 1) & 1u);
          TRACE(5);
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            st_async128_cluster(remote_recv + stage_offset<HD>(row, 2 * c), x[c], remote_recv_full);
            if (c * 16 + 8 < HD) st_async128_cluster(remote_recv + stage_offset<HD>(row, 2 * c + 1), x[c] + 8, remote_recv_full);
          }
          TRACE(7);
          ++sends;
        } else {
          // complete dQ of query block `rank`: staged in place over the receive buffer
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
#pragma unroll
            for (int e = 0; e < 16; ++e) x[c][e] *= scale;
            stage_out8<HD>(sRecv, row, 2 * c, x[c]);   // in place: this thread's own row
            if (c * 16 + 8 < HD) stage_out8<HD>(sRecv, row, 2 * c + 1, x[c] + 8);
          }
          TRACE(9);
          // dK (lanes = keys of block `rank`, like dV)
          {
            float k[kChunks][16];
#pragma unroll
            for (int c = 0; c < kChunks; ++c) tmem_ld16_nowait(tmem + lane_base + CF::kColDK + c * 16, k[c]);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bars + kDkDrained);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
#pragma unroll
              for (int e = 0; e < 16; ++e) k[c][e] *= scale;
              stage_out8<HD>(sK, row, 2 * c, k[c]);
              if (c * 16 + 8 < HD) stage_out8<HD>(sK, row, 2 * c + 1, k[c] + 8);
            }
          }
          TRACE(10);
          fence_proxy_async();
          mbar_arrive(bars + kStageFull);             // the store thread (warp 15) takes it from here
          TRACE(6);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (nst > 1) cluster_sync_all();                    // no CTA exits while a peer may still write to or signal it
  if (warp == 13) tmem_dealloc<1>(tmem, 512);
}

// delta[b, h, t] = sum_d dO[b, t, h, d] * O[b, t, h, d].  The two tensors are read as one flat stream of 16-byte chunks
// (consecutive threads = consecutive chunks: fully coalesced); a (token, head) segment is HD / 8 consecutive chunks, whose
// partial dot products meet in shared memory.
template <int HD>
__global__ void __launch_bounds__(32 * (HD / 8)) attn_delta_kernel(const uint4* __restrict__ o, const uint4* __restrict__ d_o,
                                                                   float* __restrict__ delta, int64_t segments, int T, int H) {
  constexpr int CH = HD / 8;                   // chunks per segment
  constexpr int SEG = 32;                      // segments per block
  __shared__ float part[SEG * CH];
  const int64_t seg0 = (int64_t)blockIdx.x * SEG;
  const int64_t chunk = seg0 * CH + threadIdx.x;
  part[threadIdx.x] = chunk < segments * CH ? dot8(o[chunk], d_o[chunk]) : 0.f;
  __syncthreads();
  if (threadIdx.x < SEG && seg0 + threadIdx.x < segments) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc += part[threadIdx.x * CH + c];
    const int64_t idx = seg0 + threadIdx.x;    // = token * H + head
    const int h = (int)(idx % H);
    const int64_t tok = idx / H;
    delta[((tok / T) * H + h) * T + tok % T] = acc;
  }
}

template <int HD>
int bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta, int B, int T,
               int H, cudaStream_t st) {
  static bool done = false;
  auto kernel = attn_fa_bwd_kernel<HD>;
  if (!done) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg<HD>::kTotal));
    done = true;
  }
  {
    const int64_t segments = (int64_t)B * T * H;
    attn_delta_kernel<HD><<<(unsigned)((segments + 31) / 32), 32 * (HD / 8), 0, st>>>((const uint4*)o, (const uint4*)d_o, delta,
                                                                                     segments, T, H);
    REED_LAUNCH_CHECK();
  }
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int64_t rows = (int64_t)B * T;
  if (make_map3(&maps.qkv_main, qkv, rows, 3 * H, HD, 0)) return 1;
  if (make_map3(&maps.do_main, d_o, rows, H, HD, 0)) return 1;
  if (make_map3(&maps.out_main, dqkv, rows, 3 * H, HD, 0)) return 1;
  if (Tile<HD>::kTail) {
    if (make_map3(&maps.qkv_tail, qkv, rows, 3 * H, HD, 1)) return 1;
    if (make_map3(&maps.do_tail, d_o, rows, H, HD, 1)) return 1;
    if (make_map3(&maps.out_tail8, dqkv, rows, 3 * H, HD, 2)) return 1;
  }
  const float scale = 1.f / sqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  const int nst = T / kRows;
  const int items = B * H;

  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kBwdThreads, 1, 1);
  cfg.dynamicSmemBytes = BwdCfg<HD>::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nst;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent clusters: as many as can be co-resident (1 CTA per SM; a cluster lives inside one GPC)
  static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (max_clusters[nst] == 0) {
    cfg.gridDim = dim3(nst * (sm_count() / nst), 1, 1);
    int nc = 0;
    REED_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&nc, kernel, &cfg));
    REED_REQUIRE(nc > 0, "attention backward: a cluster of %d CTAs with %d bytes of shared memory does not fit", nst,
                 BwdCfg<HD>::kTotal);
    max_clusters[nst] = nc;
  }
  const int clusters = items < max_clusters[nst] ? items : max_clusters[nst];
  cfg.gridDim = dim3(clusters * nst, 1, 1);
  REED_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, lse, (const float*)delta, T, H, items, scale, scale_log2));
  return 0;
}

}  // namespace

// cluster size = T / 128 must be a portable cluster size
bool attn_fa_bwd_supported(int T, int hd) {
  return (T == 128 || T == 256 || T == 512 || T == 1024) && (hd == 64 || hd == 72);
}

int attn_fa_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta, int B, int T,
                int H, int hd, cudaStream_t st) {
  if (B == 0) return 0;
  if (hd == 64) return bwd_launch<64>(qkv, o, d_o, lse, dqkv, delta, B, T, H, st);
  if (hd == 72) return bwd_launch<72>(qkv, o, d_o, lse, dqkv, delta, B, T, H, st);
  return fail("tcgen05 attention: head_dim %d unsupported", hd);
}

}  // namespace reed
