// HBM-bound fused kernels of the SiT block: adaLN-Zero LayerNorm+modulate (fwd/bwd), gate backward,
// casts/activations, bias-gradient column sums, token mean.  One warp owns one token row; rows are cached
// in registers (<= 16 float4 per lane, D <= 2048) so every tensor is read exactly once.
//
// Reference semantics: /root/reference/image/models/sit.py:26-27 (modulate), 113/119/146 (LayerNorm, no
// affine, eps 1e-6), 130-137 (block), 153-158 (final layer).
#include <stdlib.h>
#include "common.cuh"

namespace reed {

constexpr int kRowWarps = 4;   // warps per CTA in the row kernels

// ---------------------------------------------------------------------------------------------
// Row streaming: the HBM-bound row kernels keep `kRowStages` whole rows per CTA in flight with 1-D TMA bulk copies
// (cp.async.bulk global -> shared, completion on an mbarrier), so the bytes in flight are bounded by shared memory
// (tens of KB per CTA) instead of by registers; threads then pick their float4 column groups out of shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kRowStages = 4;

__device__ __forceinline__ uint32_t row_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void row_bar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(row_smem_u32(bar)));
}
__device__ __forceinline__ void row_bar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(row_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void row_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = row_smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// one lane of a converged warp
__device__ __forceinline__ bool row_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// bytes: multiple of 16; src, dst 16-byte aligned
__device__ __forceinline__ void row_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(row_smem_u32(dst)), "l"(src), "r"(bytes), "r"(row_smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// out[m,:] = LN(x[m,:]) * (1 + scale[g,:]) + shift[g,:],  g = m / rows_per_group
// A CTA owns `rows_per_cta` consecutive rows of ONE group: the group's shift and 1 + scale vectors are staged in shared
// memory once (they would otherwise be 2 dependent global loads per column group per row).  One warp per row at a
// time; every warp streams its rows (r0 + warp, r0 + warp + kLnWarps, ...) through a private ring of kLnStages
// shared-memory row buffers filled by bulk copies it issues itself.
// ---------------------------------------------------------------------------------------------
constexpr int kLnWarps = 8;
constexpr int kLnStages = 2;

template <typename TA, int MAXV>
__global__ void __launch_bounds__(kLnWarps * 32) ln_modulate_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale, int64_t ld_mod,
    int rows_per_group, int rows_per_cta, TA* __restrict__ out, int64_t ld_out, float* __restrict__ mean_out,
    float* __restrict__ rstd_out, int M, int D, float eps) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ uint64_t bars[kLnWarps][kLnStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, M);
  float* ring = reinterpret_cast<float*>(ln_smem) + (size_t)warp * kLnStages * D;
  float* s_shift = reinterpret_cast<float*>(ln_smem) + (size_t)kLnWarps * kLnStages * D;
  float* s_scale1 = s_shift + D;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  const int first = r0 + warp;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnStages; ++s) row_bar_init(&bars[warp][s]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();     // launched programmatically: everything above ran under the tail of the previous kernel
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnStages; ++s) {
      const int row = first + s * kLnWarps;
      if (row < r1) {
        row_bar_expect(&bars[warp][s], row_bytes);
        row_bulk_load(ring + (size_t)s * D, x + (int64_t)row * D, row_bytes, &bars[warp][s]);
      }
    }
  }
  {
    const int g = r0 / rows_per_group;
    const float* sh = shift + (int64_t)g * ld_mod;
    const float* sc = scale + (int64_t)g * ld_mod;
    for (int c = threadIdx.x * 4; c < D; c += kLnWarps * 32 * 4) {
      const F4 a = load4(sh + c);
      F4 b = load4(sc + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) b.v[j] += 1.f;
      store4(s_shift + c, a);
      store4(s_scale1 + c, b);
    }
  }
  __syncthreads();
  int it = 0;
  for (int row = first; row < r1; row += kLnWarps, ++it) {
    const int slot = it % kLnStages;
    row_bar_wait(&bars[warp][slot], (uint32_t)(it / kLnStages) & 1u);
    const float* xr = ring + (size_t)slot * D;
    F4 c[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int col = (i * 32 + lane) * 4;
      if (col < D) {
        c[i] = load4(xr + col);
        s += (c[i].v[0] + c[i].v[1]) + (c[i].v[2] + c[i].v[3]);
      }
    }
    __syncwarp();          // every lane has its values in registers: the slot can be refilled
    if (lane == 0) {
      const int nrow = row + kLnStages * kLnWarps;
      if (nrow < r1) {
        row_bar_expect(&bars[warp][slot], row_bytes);
        row_bulk_load(ring + (size_t)slot * D, x + (int64_t)nrow * D, row_bytes, &bars[warp][slot]);
      }
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int col = (i * 32 + lane) * 4;
      if (col < D) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d = c[i].v[j] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    TA* o = out + (int64_t)row * ld_out;
    if constexpr (sizeof(TA) == 2) {
      // rows with 8 spare elements carry the "ones column" that lets a weight-gradient GEMM over this matrix produce
      // the bias gradient as one more output column (reed_gemm_wgrad_bias): [1, 0, 0, 0, 0, 0, 0, 0] at column D
      if (ld_out >= D + 8 && lane == 0) *reinterpret_cast<uint4*>(o + D) = make_uint4(0x00003F80u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int col = (i * 32 + lane) * 4;
      if (col < D) {
        const F4 a = load4(s_shift + col), b = load4(s_scale1 + col);
        F4 r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = (c[i].v[j] - mean) * rstd * b.v[j] + a.v[j];
        store4(o + col, r);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-wise backward kernels of the adaLN-Zero block, one template:
//   LN   : dx[m,:] = dres[m,:] + rstd * (e - mean(e) - xhat * mean(e*xhat)),  e = dout*(1+scale[g,:])
//          dshift[g,:] += sum_m dout ; dscale[g,:] += sum_m dout*xhat
//   GATE : backward of x_new = x + gate[g,:]*y applied to the gradient just produced (LN) or to `dres` (no LN):
//          dy = gate*dx (act dtype) ; dgate[g,:] += sum_m dx*y ; dbias[:] += sum_m dy   (y = GEMM + bias)
// LN+GATE fuses the MLP-branch LayerNorm backward with the attention-branch gate backward of the same block, so the
// fp32 residual gradient dx1 is written once and never re-read by a separate gate kernel.
//
// Layout: a CTA owns `rows_per_cta` consecutive rows of ONE group and all D columns; thread t owns the float4 column
// groups t, t+NT, ... (V per thread), so every global access is a fully coalesced 16 B (8 B for bf16) per lane and the
// column partial sums live in registers for the CTA's whole row range.  Per row the only cross-thread step is the
// (sum e, sum e*xhat) block reduction.  Partials are flushed with vector red.global.add (4 floats per op).
// ---------------------------------------------------------------------------------------------
struct RowBwdParams {
  const void* dout;      // LN: [M,D] act dtype
  const float* x;        // LN: [M,D]
  const float* mean;     // LN: [M]
  const float* rstd;     // LN: [M]
  const float* scale;    // LN: rows of pitch ld_mod
  const float* dres;     // LN: optional residual gradient; !LN: the incoming gradient [M,D]
  float* dx;             // LN: [M,D]
  float* dshift;         // LN: rows of pitch ld_dmod
  float* dscale;
  const void* y;         // GATE: [M,D] act dtype
  const float* gate;     // GATE: rows of pitch ld_mod
  void* dy;              // GATE: [M,D] act dtype
  float* dgate;          // GATE: rows of pitch ld_dmod
  float* dbias;          // GATE: optional [D]
  int64_t ld_mod, ld_dmod;   // row pitch of the modulation vectors (scale, gate) / of their gradients
  int M, D, rows_per_group, rows_per_cta;
};

__device__ __forceinline__ void red_add_v4(float* p, const F4& f) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(f.v[0]), "f"(f.v[1]), "f"(f.v[2]), "f"(f.v[3])
               : "memory");
}

constexpr int kRowMaxThreads = 384;
constexpr int kRowV = 3;            // float4 column groups per thread
constexpr int kRowMaxTeams = 12;
constexpr int kRowMaxStages = 4;

// bytes of one row's operands in the shared-memory ring of row_bwd_kernel
template <typename TA, bool LN, bool GATE>
__host__ __device__ inline int row_slot_bytes(int D, bool has_res) {
  return (LN ? D * (int)sizeof(TA) + D * 4 : 0) + (has_res ? D * 4 : 0) + (GATE ? D * (int)sizeof(TA) : 0);
}

// Work split (round 2): a row is owned by a TEAM of `team_threads` threads (D / 12 rounded up to whole warps: 3 warps at
// D = 1152), every thread holding kRowV float4 column groups of it; a CTA runs 384 / team_threads teams side by side on
// interleaved rows (r0 + team, r0 + team + teams, ...), each with its own ring of `stages` row slots filled by bulk copies
// and its own named barrier.  With one float4 per thread and the whole CTA on one row (round 1) the per-row overhead
// - two shuffle reductions, a CTA barrier, the cross-warp sum, address arithmetic - was 4/5 of the 256 instructions a
// warp issued per row and the kernels ran at 0.56-0.67 issue utilisation, 0.65-0.73 of the HBM rate (ncu,
// profiles/r02_ncu_row_bwd.txt); twelve elements per thread amortise it.
// FULL: D == 4 * V * team_threads, every thread owns V valid column groups - no per-group predicates (at D = 1152 they and
// the divergence bookkeeping around them were a fifth of the 576 instructions a warp issued per row).
template <typename TA, int V, bool LN, bool GATE, bool FULL>
__global__ void __launch_bounds__(kRowMaxThreads, 1) row_bwd_kernel(const RowBwdParams p, int team_threads, int stages) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  extern __shared__ __align__(128) uint8_t row_smem[];
  __shared__ uint64_t full[kRowMaxTeams * kRowMaxStages];
  __shared__ float2 red[2][kRowMaxThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int teams = blockDim.x / team_threads, team = tid / team_threads, t = tid - team * team_threads;
  const int wpt = team_threads >> 5, warp0 = team * wpt;
  const int D = p.D;
  const int r0 = blockIdx.x * p.rows_per_cta;
  const int r1 = min(r0 + p.rows_per_cta, p.M);
  TA* __restrict__ dy = reinterpret_cast<TA*>(p.dy);
  const float inv_d = 1.f / D;
  const bool has_res = !LN || p.dres != nullptr;
  // slot layout: [dout TA | x f32] (LN) [dres f32] (has_res) [y TA] (GATE)
  const int o_x = LN ? D * (int)sizeof(TA) : 0;
  const int o_r = o_x + (LN ? D * 4 : 0);
  const int o_y = o_r + (has_res ? D * 4 : 0);
  const int slot_bytes = o_y + (GATE ? D * (int)sizeof(TA) : 0);
  uint8_t* ring = row_smem + (size_t)team * stages * slot_bytes;
  uint64_t* bars = full + team * kRowMaxStages;

  auto issue = [&](int row, int slot) {       // team thread 0: all operands of one row -> ring slot
    uint8_t* dst = ring + (size_t)slot * slot_bytes;
    const int64_t base = (int64_t)row * D;
    row_bar_expect(&bars[slot], (uint32_t)slot_bytes);
    if constexpr (LN) {
      row_bulk_load(dst, reinterpret_cast<const TA*>(p.dout) + base, (uint32_t)(D * sizeof(TA)), &bars[slot]);
      row_bulk_load(dst + o_x, p.x + base, (uint32_t)D * 4u, &bars[slot]);
    }
    if (has_res) row_bulk_load(dst + o_r, p.dres + base, (uint32_t)D * 4u, &bars[slot]);
    if constexpr (GATE) row_bulk_load(dst + o_y, reinterpret_cast<const TA*>(p.y) + base, (uint32_t)(D * sizeof(TA)), &bars[slot]);
  };
  auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(team_threads) : "memory"); };
  const int first = r0 + team;
  if (t == 0) {
    for (int s = 0; s < stages; ++s) row_bar_init(&bars[s]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();     // launched programmatically: everything above ran under the tail of the previous kernel
  if (t == 0) {
    for (int s = 0; s < stages; ++s)
      if (first + s * teams < r1) issue(first + s * teams, s);
  }
  __syncthreads();

  int col[V];
  bool ok[V];
  F4 one_plus[LN ? V : 1], gt[GATE ? V : 1];
  F4 a_sh[LN ? V : 1], a_sc[LN ? V : 1], a_g[GATE ? V : 1], a_b[GATE ? V : 1];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    col[i] = (t + i * team_threads) * 4;
    ok[i] = FULL || col[i] < D;
  }
  auto load_group = [&](int g) {              // per-group modulation vectors; column partial sums restart
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if constexpr (LN) {
        a_sh[i] = F4{{0, 0, 0, 0}};
        a_sc[i] = F4{{0, 0, 0, 0}};
        one_plus[i] = F4{{1, 1, 1, 1}};
        if (ok[i]) {
          const F4 s = load4(p.scale + (int64_t)g * p.ld_mod + col[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) one_plus[i].v[j] = 1.f + s.v[j];
        }
      }
      if constexpr (GATE) {
        a_g[i] = F4{{0, 0, 0, 0}};
        a_b[i] = F4{{0, 0, 0, 0}};
        gt[i] = F4{{0, 0, 0, 0}};
        if (ok[i]) gt[i] = load4(p.gate + (int64_t)g * p.ld_mod + col[i]);
      }
    }
  };
  auto flush_group = [&](int g) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if (ok[i]) {
        const int64_t off = (int64_t)g * p.ld_dmod + col[i];
        if constexpr (LN) {
          red_add_v4(p.dshift + off, a_sh[i]);
          red_add_v4(p.dscale + off, a_sc[i]);
        }
        if constexpr (GATE) {
          red_add_v4(p.dgate + off, a_g[i]);
          if (p.dbias != nullptr) red_add_v4(p.dbias + col[i], a_b[i]);
        }
      }
    }
  };

  if (first >= r1) return;                    // (after the CTA barrier; a team without rows owns no named barrier traffic)
  int g = first / p.rows_per_group;
  int g_end = (g + 1) * p.rows_per_group;     // first row of the next group
  load_group(g);
  float mean_n = 0.f, rstd_n = 0.f;
  if constexpr (LN) { mean_n = p.mean[first]; rstd_n = p.rstd[first]; }
  int buf = 0, slot = 0;
  uint32_t parity = 0;
  for (int row = first; row < r1; row += teams) {
    if (row >= g_end) {                       // the team's rows cross into the next sample (rows_per_group >= teams or not)
      flush_group(g);
      g = row / p.rows_per_group;
      g_end = (g + 1) * p.rows_per_group;
      load_group(g);
    }
    const float mean = mean_n, rstd = rstd_n;
    if constexpr (LN) {
      if (row + teams < r1) { mean_n = p.mean[row + teams]; rstd_n = p.rstd[row + teams]; }
    }
    row_bar_wait(&bars[slot], parity);
    const uint8_t* src = ring + (size_t)slot * slot_bytes;
    F4 d[LN ? V : 1], xv[LN ? V : 1], gin[V], yv[GATE ? V : 1];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      gin[i] = F4{{0, 0, 0, 0}};
      if (ok[i]) {
        if constexpr (LN) {
          d[i] = load4(reinterpret_cast<const TA*>(src) + col[i]);
          xv[i] = load4(reinterpret_cast<const float*>(src + o_x) + col[i]);
        }
        if (has_res) gin[i] = load4(reinterpret_cast<const float*>(src + o_r) + col[i]);
        if constexpr (GATE) yv[i] = load4(reinterpret_cast<const TA*>(src + o_y) + col[i]);
      }
    }
    float s1 = 0.f, s2 = 0.f;
    if constexpr (LN) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (ok[i]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h = (xv[i].v[j] - mean) * rstd;
            const float dv = d[i].v[j];
            xv[i].v[j] = h;
            a_sh[i].v[j] += dv;
            a_sc[i].v[j] += dv * h;
            const float e = dv * one_plus[i].v[j];
            d[i].v[j] = e;
            s1 += e;
            s2 += e * h;
          }
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (wpt > 1 && lane == 0) red[buf][warp] = make_float2(s1, s2);
    }
    // one team barrier per row: every thread of the team has copied its operands out of the slot (it can be refilled)
    // and, for LN, the per-warp partial sums are visible
    team_sync();
    // the team's first warp refills the slot: a warp-uniform branch and an elected lane, so the bulk-copy operands sit in
    // uniform registers (issued from a one-lane divergent branch each copy is wrapped in an R2UR waterfall loop)
    if (t < 32 && row + stages * teams < r1) {
      if (row_elect_one()) issue(row + stages * teams, slot);
      __syncwarp();
    }
    if (++slot == stages) { slot = 0; parity ^= 1u; }
    const int64_t base = (int64_t)row * D;
    if constexpr (LN) {
      if (wpt > 1) {
        s1 = 0.f;
        s2 = 0.f;
        for (int w = 0; w < wpt; ++w) {
          const float2 tt = red[buf][warp0 + w];
          s1 += tt.x;
          s2 += tt.y;
        }
        buf ^= 1;    // the next row uses the other slot of `red`
      }
      const float c1 = s1 * inv_d, c2 = s2 * inv_d;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (ok[i]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) gin[i].v[j] += rstd * (d[i].v[j] - c1 - xv[i].v[j] * c2);
          store4(p.dx + base + col[i], gin[i]);
        }
      }
    }
    if constexpr (GATE) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (ok[i]) {
          F4 o;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a_g[i].v[j] += gin[i].v[j] * yv[i].v[j];
            o.v[j] = to_f(from_f<TA>(gin[i].v[j] * gt[i].v[j]));
            a_b[i].v[j] += o.v[j];
          }
          store4(dy + base + col[i], o);
        }
      }
    }
  }
  flush_group(g);
}

// ---------------------------------------------------------------------------------------------
// column sums: out[n] (+)= sum_m src[m, n]      (bias gradients)
// ---------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256) colsum_kernel(const TA* __restrict__ src, int64_t ld, float* __restrict__ out,
                                                      int M, int N, int rows_per_cta) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, M);
  F4 acc{{0, 0, 0, 0}};
  if (col < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      F4 v = load4(src + (int64_t)r * ld + col);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc.v[j] += v.v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[warp][lane * 4 + j] = acc.v[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    int c = blockIdx.x * 128 + threadIdx.x;
    if (c < N) atomicAdd(out + c, s);
  }
}

// ---------------------------------------------------------------------------------------------
// flat elementwise: casts, SiLU fwd/bwd, act-derivative multiply
// ---------------------------------------------------------------------------------------------
enum UnaryOp { kCast = 0, kSilu = 1, kGeluErf = 2 };   // kGeluErf: nn.GELU() exact, the DINOv2 target encoders' MLP

template <typename TI, typename TO, int OP>
__global__ void __launch_bounds__(256) unary_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n4) {
  pdl_launch();   // dependents (the next GEMM of the stream) may start their prologue
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 v = load4(in + i * 4);
    if (OP == kSilu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v.v[j] = silu(v.v[j]);
    }
    if (OP == kGeluErf) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v.v[j] = 0.5f * v.v[j] * (1.f + erff(v.v[j] * 0.70710678118654752f));
    }
    store4(out + i * 4, v);
  }
}

// dx = dy * f'(h), f in {silu, gelu_tanh}; TH = dtype of h, TD = dtype of dy/dx
template <typename TH, typename TD, int GELU>
__global__ void __launch_bounds__(256) act_bwd_kernel(const TD* __restrict__ dy, const TH* __restrict__ h,
                                                       TD* __restrict__ dx, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 d = load4(dy + i * 4), hv = load4(h + i * 4), o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = d.v[j] * (GELU ? gelu_tanh_grad(hv.v[j]) : silu_grad(hv.v[j]));
    store4(dx + i * 4, o);
  }
}

// out[g, :] = mean over the group's rows of x (token mean for the 't' projector); bwd broadcasts dy/T
template <typename TO>
__global__ void __launch_bounds__(256) group_mean_kernel(const float* __restrict__ x, TO* __restrict__ out,
                                                          int rows_per_group, int D) {
  const int g = blockIdx.y;
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (col >= D) return;
  F4 acc{{0, 0, 0, 0}};
  const float* base = x + (int64_t)g * rows_per_group * D + col;
  for (int r = 0; r < rows_per_group; ++r) {
    F4 v = load4(base + (int64_t)r * D);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.v[j] += v.v[j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) acc.v[j] /= rows_per_group;
  store4(out + (int64_t)g * D + col, acc);
}

// dx[m,:] (+)= dy[g,:] / T
__global__ void __launch_bounds__(256) group_mean_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                              int rows_per_group, int D, int64_t n4, int accumulate) {
  const int d4 = D / 4;
  const float inv = 1.f / rows_per_group;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / d4;
    int col = (int)(i - row * d4) * 4;
    F4 v = load4(dy + (row / rows_per_group) * D + col);
    F4 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = v.v[j] * inv;
    if (accumulate) {
      F4 p = load4(dx + i * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) o.v[j] += p.v[j];
    }
    store4(dx + i * 4, o);
  }
}

// out = a + b (fp32), used to merge the projector-tap gradient into the residual-stream gradient
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   float* __restrict__ out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    F4 x = load4(a + i * 4), y = load4(b + i * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) x.v[j] += y.v[j];
    store4(out + i * 4, x);
  }
}

static inline int flat_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

static int row_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
  }
  return n;
}

// Launch with the programmatic-serialization attribute: the grid may become resident while the previous kernel of the
// stream drains (its prologue - barrier init, index arithmetic - runs there); the kernel calls pdl_wait() before its
// first global access (A/B on one box: 974 -> 978 img/s).
template <typename... KArgs, typename... Args>
static cudaError_t launch_row_kernel(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename TA, int V>
static int ln_fwd_launch(const float* x, const float* shift, const float* scale, int64_t ld_mod, int rpg, void* out,
                         int64_t ld_out, float* mean, float* rstd, int M, int D, float eps, cudaStream_t st) {
  auto kernel = ln_modulate_fwd_kernel<TA, V>;
  const int smem = (kLnWarps * kLnStages + 2) * D * 4;
  static int configured = 0;
  if (configured < smem) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  // row ranges never straddle a group: the largest power of two <= 32 that divides rows_per_group, shrunk until the
  // grid covers every SM (measured at M = 8192, D = 1152: 32 rows per CTA = 256 CTAs in one wave 16.4 us, 16 rows = 512
  // CTAs in 1.7 waves 18.4 us, 64 rows 18.5 us)
  int rows = 32;
  while (rows > 1 && (rpg % rows != 0 || ceil_div(M, rows) < row_sm_count())) rows >>= 1;
  REED_CHECK_CUDA(launch_row_kernel(kernel, ceil_div(M, rows), kLnWarps * 32, smem, st, x, shift, scale, ld_mod, rpg, rows, (TA*)out, ld_out,
                                    mean, rstd, M, D, eps));
  return 0;
}

template <typename TA>
static int ln_fwd_dispatch(const float* x, const float* shift, const float* scale, int64_t ld_mod, int rpg, void* out,
                           int64_t ld_out, float* mean, float* rstd, int M, int D, float eps, cudaStream_t st) {
  int nv = ceil_div(D, 128);
#define LN_FWD(V) return ln_fwd_launch<TA, V>(x, shift, scale, ld_mod, rpg, out, ld_out, mean, rstd, M, D, eps, st)
  if (nv <= 4) LN_FWD(4); else if (nv <= 9) LN_FWD(9); else if (nv <= 12) LN_FWD(12); else LN_FWD(16);
#undef LN_FWD
}

template <typename TA, bool LN, bool GATE>
static int row_bwd_dispatch(RowBwdParams p, cudaStream_t st) {
  const int D = p.D;
  const bool has_res = !LN || p.dres != nullptr;
  // a team = the threads of one row: kRowV float4 column groups each, whole warps (3 warps at D = 1152, 1 at D = 384)
  const int team_threads = ceil_div(ceil_div(D / 4, kRowV), 32) * 32;
  const bool full = D == 4 * kRowV * team_threads;
  auto kernel = full ? row_bwd_kernel<TA, kRowV, LN, GATE, true> : row_bwd_kernel<TA, kRowV, LN, GATE, false>;
  REED_REQUIRE(team_threads <= kRowMaxThreads, "row kernels: D = %d exceeds %d columns", D, kRowMaxThreads * kRowV * 4);
  int teams = kRowMaxThreads / team_threads;
  if (teams > kRowMaxTeams) teams = kRowMaxTeams;
  // one CTA per SM (twelve elements per thread need the registers); the ring takes the shared memory: as many row slots
  // per team as fit, at most four
  const int slot = row_slot_bytes<TA, LN, GATE>(D, has_res);
  int stages = (200 * 1024) / (teams * slot);
  if (stages > kRowMaxStages) stages = kRowMaxStages;
  while (stages < 2 && teams > 1) { --teams; stages = (200 * 1024) / (teams * slot); if (stages > kRowMaxStages) stages = kRowMaxStages; }
  REED_REQUIRE(stages >= 1, "row kernels: a row of D = %d does not fit the shared-memory ring", D);
  const int smem = teams * stages * slot;
  static int configured[2] = {0, 0};     // per template instance (FULL or not)
  if (configured[full] < smem) {
    REED_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[full] = smem;
  }
  // one balanced wave: equal row ranges that may cross sample boundaries (a team flushes its column partials when the
  // group changes)
  int rows = ceil_div(p.M, row_sm_count());
  if (rows < teams) rows = teams;
  p.rows_per_cta = rows;
  REED_CHECK_CUDA(launch_row_kernel(kernel, ceil_div(p.M, rows), teams * team_threads, smem, st, p, team_threads, stages));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Optional per-head LayerNorm of q and k (timm Attention(qk_norm=True): nn.LayerNorm(head_dim), affine, eps 1e-5),
// applied to the packed qkv [rows, 3, H, hd]; v is copied through.  One warp per (row, head) unit of one of q/k/v
// (blockIdx.y), element i of the unit in lane i % 32.  Backward keeps dweight/dbias partials in registers across
// the warp's units and flushes them with atomics.
// ---------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256) qk_norm_fwd_kernel(const TA* __restrict__ qkv, const float* __restrict__ wq,
                                                           const float* __restrict__ bq, const float* __restrict__ wk,
                                                           const float* __restrict__ bk, TA* __restrict__ out,
                                                           float* __restrict__ stats, int64_t rows, int H, int hd, float eps) {
  const int lane = threadIdx.x & 31;
  const int which = blockIdx.y;
  const int64_t units = rows * H;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float* w = which == 0 ? wq : wk;
  const float* b = which == 0 ? bq : bk;
  for (int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += warps) {
    const int64_t row = u / H;
    const int h = (int)(u - row * H);
    const int64_t base = (row * 3 + which) * (int64_t)H * hd + (int64_t)h * hd;
    float v[4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < hd ? to_f(qkv[base + c]) : 0.f;
      s += v[i];
    }
    if (which == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < hd) out[base + c] = from_f<TA>(v[i]);
      }
      continue;
    }
    const float mean = warp_sum(s) / hd;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < hd) q += (v[i] - mean) * (v[i] - mean);
    }
    const float rstd = rsqrtf(warp_sum(q) / hd + eps);
    if (lane == 0) {
      stats[(u * 2 + which) * 2] = mean;
      stats[(u * 2 + which) * 2 + 1] = rstd;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < hd) out[base + c] = from_f<TA>((v[i] - mean) * rstd * w[c] + b[c]);
    }
  }
}

template <typename TA>
__global__ void __launch_bounds__(256) qk_norm_bwd_kernel(const TA* __restrict__ dout, const TA* __restrict__ qkv,
                                                           const float* __restrict__ stats, const float* __restrict__ wq,
                                                           const float* __restrict__ wk, TA* __restrict__ dqkv,
                                                           float* __restrict__ dwq, float* __restrict__ dbq,
                                                           float* __restrict__ dwk, float* __restrict__ dbk, int64_t rows,
                                                           int H, int hd) {
  const int lane = threadIdx.x & 31;
  const int which = blockIdx.y;
  const int64_t units = rows * H;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float* w = which == 0 ? wq : wk;
  float wv[4], dw[4], db[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + 32 * i;
    wv[i] = (which < 2 && c < hd) ? w[c] : 0.f;
    dw[i] = 0.f;
    db[i] = 0.f;
  }
  for (int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += warps) {
    const int64_t row = u / H;
    const int h = (int)(u - row * H);
    const int64_t base = (row * 3 + which) * (int64_t)H * hd + (int64_t)h * hd;
    float dy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      dy[i] = c < hd ? to_f(dout[base + c]) : 0.f;
    }
    if (which == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < hd) dqkv[base + c] = from_f<TA>(dy[i]);
      }
      continue;
    }
    const float mean = stats[(u * 2 + which) * 2], rstd = stats[(u * 2 + which) * 2 + 1];
    float xh[4], g[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      xh[i] = c < hd ? (to_f(qkv[base + c]) - mean) * rstd : 0.f;
      g[i] = dy[i] * wv[i];
      s1 += g[i];
      s2 += g[i] * xh[i];
      dw[i] += dy[i] * xh[i];
      db[i] += dy[i];
    }
    const float c1 = warp_sum(s1) / hd, c2 = warp_sum(s2) / hd;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < hd) dqkv[base + c] = from_f<TA>(rstd * (g[i] - c1 - xh[i] * c2));
    }
  }
  if (which < 2) {
    float* dwp = which == 0 ? dwq : dwk;
    float* dbp = which == 0 ? dbq : dbk;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < hd) {
        atomicAdd(dwp + c, dw[i]);
        atomicAdd(dbp + c, db[i]);
      }
    }
  }
}

}  // namespace reed

using namespace reed;

#define ROW_ARGS_OK(M, D, rpg)                                                                   \
  REED_REQUIRE((D) % 8 == 0 && (D) <= 2048, "row kernels need D %% 8 == 0 and D <= 2048, got %d", (int)(D)); \
  REED_REQUIRE((rpg) > 0 && (M) % (rpg) == 0, "M=%d not a multiple of rows_per_group=%d", (int)(M), (int)(rpg))

extern "C" int reed_ln_modulate_fwd(const void* x, const void* shift, const void* scale, int64_t ld_mod,
                                    int rows_per_group, void* out, int64_t ld_out, int act_dtype, void* mean, void* rstd,
                                    int M, int D, float eps, void* stream) {
  ROW_ARGS_OK(M, D, rows_per_group);
  REED_REQUIRE(ld_out >= D && ld_out % 8 == 0, "ln_modulate_fwd: output pitch must be >= D and a multiple of 8 elements");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16)
    return ln_fwd_dispatch<bf16>((const float*)x, (const float*)shift, (const float*)scale, ld_mod, rows_per_group, out,
                                 ld_out, (float*)mean, (float*)rstd, M, D, eps, st);
  return ln_fwd_dispatch<float>((const float*)x, (const float*)shift, (const float*)scale, ld_mod, rows_per_group, out,
                                ld_out, (float*)mean, (float*)rstd, M, D, eps, st);
}

#define ROW_MOD_OK(ld) REED_REQUIRE((ld) % 4 == 0, "row kernels need modulation rows with a pitch that is a multiple of 4 floats")

extern "C" int reed_ln_modulate_bwd(const void* dout, int act_dtype, const void* x, const void* mean, const void* rstd,
                                    const void* scale, int64_t ld_mod, int64_t ld_dmod, int rows_per_group, const void* dres,
                                    void* dx, void* dshift, void* dscale, int M, int D, void* stream) {
  ROW_ARGS_OK(M, D, rows_per_group);
  ROW_MOD_OK(ld_mod);
  ROW_MOD_OK(ld_dmod);
  if (M == 0) return 0;
  RowBwdParams p{};
  p.dout = dout; p.x = (const float*)x; p.mean = (const float*)mean; p.rstd = (const float*)rstd;
  p.scale = (const float*)scale; p.dres = (const float*)dres; p.dx = (float*)dx; p.dshift = (float*)dshift;
  p.dscale = (float*)dscale; p.ld_mod = ld_mod; p.ld_dmod = ld_dmod; p.M = M; p.D = D; p.rows_per_group = rows_per_group;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16) return row_bwd_dispatch<bf16, true, false>(p, st);
  return row_bwd_dispatch<float, true, false>(p, st);
}

extern "C" int reed_ln_modulate_gate_bwd(const void* dout, int act_dtype, const void* x, const void* mean,
                                         const void* rstd, const void* scale, int64_t ld_mod, int64_t ld_dmod,
                                         int rows_per_group, const void* dres, void* dx, void* dshift, void* dscale,
                                         const void* y, const void* gate, void* dy, void* dgate, void* dbias, int M, int D,
                                         void* stream) {
  ROW_ARGS_OK(M, D, rows_per_group);
  ROW_MOD_OK(ld_mod);
  ROW_MOD_OK(ld_dmod);
  if (M == 0) return 0;
  RowBwdParams p{};
  p.dout = dout; p.x = (const float*)x; p.mean = (const float*)mean; p.rstd = (const float*)rstd;
  p.scale = (const float*)scale; p.dres = (const float*)dres; p.dx = (float*)dx; p.dshift = (float*)dshift;
  p.dscale = (float*)dscale; p.y = y; p.gate = (const float*)gate; p.dy = dy; p.dgate = (float*)dgate;
  p.dbias = (float*)dbias; p.ld_mod = ld_mod; p.ld_dmod = ld_dmod; p.M = M; p.D = D; p.rows_per_group = rows_per_group;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16) return row_bwd_dispatch<bf16, true, true>(p, st);
  return row_bwd_dispatch<float, true, true>(p, st);
}

extern "C" int reed_gate_bwd(const void* dxn, const void* y, int act_dtype, const void* gate, int64_t ld_mod, int64_t ld_dmod,
                             int rows_per_group, void* dy, void* dgate, void* dbias, int M, int D, void* stream) {
  ROW_ARGS_OK(M, D, rows_per_group);
  ROW_MOD_OK(ld_mod);
  ROW_MOD_OK(ld_dmod);
  if (M == 0) return 0;
  RowBwdParams p{};
  p.dres = (const float*)dxn; p.y = y; p.gate = (const float*)gate; p.dy = dy; p.dgate = (float*)dgate;
  p.dbias = (float*)dbias; p.ld_mod = ld_mod; p.ld_dmod = ld_dmod; p.M = M; p.D = D; p.rows_per_group = rows_per_group;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16) return row_bwd_dispatch<bf16, false, true>(p, st);
  return row_bwd_dispatch<float, false, true>(p, st);
}

extern "C" int reed_colsum(const void* src, int act_dtype, int64_t ld, void* out, int M, int N, void* stream) {
  REED_REQUIRE(N % 4 == 0 && ld % 4 == 0, "colsum needs N, ld %% 4 == 0");
  if (M == 0 || N == 0) return 0;
  int rows_per_cta = 256;
  dim3 grid(ceil_div(N, 128), ceil_div(M, rows_per_cta));
  if (act_dtype == kBF16)
    colsum_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, ld, (float*)out, M, N, rows_per_cta);
  else
    colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, ld, (float*)out, M, N, rows_per_cta);
  REED_LAUNCH_CHECK();
  return 0;
}

// op: 0 cast, 1 silu, 2 gelu (erf).  in_dtype/out_dtype: 0 fp32, 1 bf16
extern "C" int reed_unary(const void* in, int in_dtype, void* out, int out_dtype, int op, int64_t n, void* stream) {
  REED_REQUIRE(n % 4 == 0, "unary needs n %% 4 == 0, got %lld", (long long)n);
  if (n == 0) return 0;
  int64_t n4 = n / 4;
  int grid = flat_grid(n4);
  cudaStream_t st = (cudaStream_t)stream;
  REED_REQUIRE(op >= kCast && op <= kGeluErf, "reed_unary: unknown op %d", op);
#define UN(TI, TO, OP) unary_kernel<TI, TO, OP><<<grid, 256, 0, st>>>((const TI*)in, (TO*)out, n4)
#define UN3(TI, TO) do { if (op == kSilu) UN(TI, TO, kSilu); else if (op == kGeluErf) UN(TI, TO, kGeluErf); else UN(TI, TO, kCast); } while (0)
  if (in_dtype == kF32 && out_dtype == kF32) UN3(float, float);
  else if (in_dtype == kF32 && out_dtype == kBF16) UN3(float, bf16);
  else if (in_dtype == kBF16 && out_dtype == kF32) UN3(bf16, float);
  else if (in_dtype == kBF16 && out_dtype == kBF16) UN3(bf16, bf16);
  else return fail("reed_unary: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
#undef UN3
#undef UN
  REED_LAUNCH_CHECK();
  return 0;
}

// dx = dy * act'(h); act: 1 gelu_tanh, 2 silu.  h_dtype: dtype of h; d_dtype: dtype of dy/dx
extern "C" int reed_act_bwd(const void* dy, int d_dtype, const void* h, int h_dtype, void* dx, int act, int64_t n,
                            void* stream) {
  REED_REQUIRE(n % 4 == 0, "act_bwd needs n %% 4 == 0");
  if (n == 0) return 0;
  int64_t n4 = n / 4;
  int grid = flat_grid(n4);
  cudaStream_t st = (cudaStream_t)stream;
  const bool gelu = act == 1;
#define AB(TH, TD) do { if (gelu) act_bwd_kernel<TH, TD, 1><<<grid, 256, 0, st>>>((const TD*)dy, (const TH*)h, (TD*)dx, n4); \
                        else act_bwd_kernel<TH, TD, 0><<<grid, 256, 0, st>>>((const TD*)dy, (const TH*)h, (TD*)dx, n4); } while (0)
  if (h_dtype == kF32 && d_dtype == kF32) AB(float, float);
  else if (h_dtype == kBF16 && d_dtype == kBF16) AB(bf16, bf16);
  else if (h_dtype == kF32 && d_dtype == kBF16) AB(float, bf16);
  else AB(bf16, float);
#undef AB
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_group_mean_fwd(const void* x, void* out, int out_dtype, int groups, int rows_per_group, int D,
                                   void* stream) {
  REED_REQUIRE(D % 4 == 0, "group_mean needs D %% 4 == 0");
  if (groups == 0) return 0;
  dim3 grid(ceil_div(D / 4, 256), groups);
  if (out_dtype == kBF16)
    group_mean_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (bf16*)out, rows_per_group, D);
  else
    group_mean_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)out, rows_per_group, D);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_group_mean_bwd(const void* dy, void* dx, int groups, int rows_per_group, int D, int accumulate,
                                   void* stream) {
  REED_REQUIRE(D % 4 == 0, "group_mean needs D %% 4 == 0");
  int64_t n4 = (int64_t)groups * rows_per_group * D / 4;
  if (n4 == 0) return 0;
  group_mean_bwd_kernel<<<flat_grid(n4), 256, 0, (cudaStream_t)stream>>>((const float*)dy, (float*)dx, rows_per_group, D,
                                                                        n4, accumulate);
  REED_LAUNCH_CHECK();
  return 0;
}

extern "C" int reed_add_f32(const void* a, const void* b, void* out, int64_t n, void* stream) {
  REED_REQUIRE(n % 4 == 0, "add needs n %% 4 == 0");
  if (n == 0) return 0;
  add_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float*)a, (const float*)b, (float*)out, n / 4);
  REED_LAUNCH_CHECK();
  return 0;
}

// q/k LayerNorm over head_dim of the packed qkv (timm Attention q_norm / k_norm, affine, eps 1e-5); stats [rows,H,2,2]
extern "C" int reed_qk_norm_fwd(const void* qkv, int act_dtype, const void* wq, const void* bq, const void* wk,
                                const void* bk, void* out, void* stats, int64_t rows, int H, int hd, float eps,
                                void* stream) {
  REED_REQUIRE(hd >= 1 && hd <= 128, "qk_norm: head_dim %d unsupported (1..128)", hd);
  if (rows == 0) return 0;
  int64_t blocks = (rows * H + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  dim3 grid((unsigned)blocks, 3);
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16)
    qk_norm_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)qkv, (const float*)wq, (const float*)bq, (const float*)wk,
                                                    (const float*)bk, (bf16*)out, (float*)stats, rows, H, hd, eps);
  else
    qk_norm_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)qkv, (const float*)wq, (const float*)bq, (const float*)wk,
                                                     (const float*)bk, (float*)out, (float*)stats, rows, H, hd, eps);
  REED_LAUNCH_CHECK();
  return 0;
}

// dqkv (raw) from the gradient w.r.t. the normalised qkv; dwq/dbq/dwk/dbk fp32 [hd] are accumulated into
extern "C" int reed_qk_norm_bwd(const void* dout, int act_dtype, const void* qkv, const void* stats, const void* wq,
                                const void* wk, void* dqkv, void* dwq, void* dbq, void* dwk, void* dbk, int64_t rows,
                                int H, int hd, void* stream) {
  REED_REQUIRE(hd >= 1 && hd <= 128, "qk_norm: head_dim %d unsupported (1..128)", hd);
  if (rows == 0) return 0;
  int64_t blocks = (rows * H + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  dim3 grid((unsigned)blocks, 3);
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == kBF16)
    qk_norm_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dout, (const bf16*)qkv, (const float*)stats,
                                                    (const float*)wq, (const float*)wk, (bf16*)dqkv, (float*)dwq,
                                                    (float*)dbq, (float*)dwk, (float*)dbk, rows, H, hd);
  else
    qk_norm_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dout, (const float*)qkv, (const float*)stats,
                                                     (const float*)wq, (const float*)wk, (float*)dqkv, (float*)dwq,
                                                     (float*)dbq, (float*)dwk, (float*)dbk, rows, H, hd);
  REED_LAUNCH_CHECK();
  return 0;
}
