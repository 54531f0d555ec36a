// Probe (round 2): tcgen05.mma with the A operand in TENSOR MEMORY - the form the rebuilt attention kernels use for
// O += P V (P written by the softmax warps with tcgen05.st, never touching shared memory).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I reed_b200/csrc profiles/probe_ts_mma.cu -o /tmp/probe_ts
//
// Checks, against a host reference: (1) the layout assumption - a bf16 A[128 x K] lives in TMEM as lane = row, 32-bit
// column j = elements (2j, 2j+1), low half first, 8 columns per K = 16 step; (2) B = a [keys x 64] SWIZZLE_128B tile read
// MN-major (the V tile exactly as TMA stages it) and the 16-wide SWIZZLE_32B tail; (3) the one-hot "which k is this
// slot" table that would reveal a different mapping.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "tcgen05_ptx.cuh"

namespace reed {
thread_local char g_err[512];
int fail(const char*, ...) { return 1; }
}  // namespace reed
using namespace reed;

constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) { return (sbo_bytes >> 4) | (1u << 14) | (layout << 29); }
constexpr uint32_t kHiSw128 = desc_hi(1024, 2);
constexpr uint32_t kHiSw32 = desc_hi(256, 6);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// p: [128][128] bf16 row-major (A), v: [128 keys][80] bf16 row-major (cols 72..79 zero), d: [128][80] fp32
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* p, const __nv_bfloat16* v, float* d) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sV = smem;                    // [128 x 64] SWIZZLE_128B (16 KB) + [128 x 16] SWIZZLE_32B tail (4 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20480);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // stage V the way TMA would: row r, 16-byte chunk c at (c ^ (r & 7)); tail rows of 32 B, chunk c at c ^ ((r >> 2) & 1)
  for (int i = tid; i < 128 * 10; i += 128) {
    const int r = i / 10, c = i % 10;
    const uint4 val = *reinterpret_cast<const uint4*>(v + r * 80 + c * 8);
    uint8_t* dst = c < 8 ? sV + r * 128 + ((c ^ (r & 7)) << 4) : sV + 16384 + r * 32 + (((c - 8) ^ ((r >> 2) & 1)) << 4);
    *reinterpret_cast<uint4*>(dst) = val;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<1>(tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  // P row of this thread -> TMEM columns 128..191 (bf16 pairs)
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = *reinterpret_cast<const uint32_t*>(p + tid * 128 + 2 * (c0 + j));
    tmem_st16(trow + 128 + c0, r);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc64 = make_idesc(128, 64, 0, 1);
    constexpr uint32_t idesc16 = make_idesc(128, 16, 0, 1);
    const uint32_t lz = desc_lo(smem_u32(sV));
    for (int ks = 0; ks < 8; ++ks) {
      umma_ts(tmem, tmem + 128 + ks * 8, mk_desc(lz + ks * (2048 >> 4), kHiSw128), idesc64, ks > 0);
      umma_ts(tmem + 64, tmem + 128 + ks * 8, mk_desc(lz + (16384 >> 4) + ks * (512 >> 4), kHiSw32), idesc16, ks > 0);
    }
    umma_commit<1>(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 96; c0 += 32) {
    float x[32];
    tmem_ld32(trow + c0, x);
    for (int j = 0; j < 32; ++j)
      if (c0 + j < 80) d[tid * 80 + c0 + j] = x[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<__nv_bfloat16> hp(128 * 128), hv(128 * 80);
  std::vector<float> fp(128 * 128), fv(128 * 80), hd(128 * 80), ref(128 * 80);
  __nv_bfloat16 *dp, *dv;
  float* dd;
  cudaMalloc(&dp, hp.size() * 2);
  cudaMalloc(&dv, hv.size() * 2);
  cudaMalloc(&dd, hd.size() * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576);
  int bad_total = 0;
  for (int test = 0; test < 10; ++test) {
    srand(test + 1);
    const int onehot[9] = {0, 1, 2, 15, 16, 17, 64, 65, 127};
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 128; ++k) {
        float x = test == 0 ? (rand() % 2001 - 1000) / 1000.f : (k == onehot[test - 1] ? 1.f : 0.f);
        fp[m * 128 + k] = bf(x);
        hp[m * 128 + k] = __float2bfloat16(x);
      }
    for (int k = 0; k < 128; ++k)
      for (int n = 0; n < 80; ++n) {
        float x = n >= 72 ? 0.f : (test == 0 ? (rand() % 2001 - 1000) / 1000.f : (float)k + (n == 71 ? 0.5f : 0.f));
        fv[k * 80 + n] = bf(x);
        hv[k * 80 + n] = __float2bfloat16(x);
      }
    cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, hv.data(), hv.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, hd.size() * 4);
    probe_kernel<<<1, 128, 24576>>>(dp, dv, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("test %d: CUDA error %s\n", test, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 80; ++n) {
        double acc = 0;
        for (int k = 0; k < 128; ++k) acc += (double)fp[m * 128 + k] * fv[k * 80 + n];
        ref[m * 80 + n] = (float)acc;
        const double err = fabs(acc - hd[m * 80 + n]);
        if (err > worst) worst = err;
        if (err > 1e-2) ++bad;
      }
    if (test == 0)
      printf("random P.V: max |err| %.3e, %d of %d elements off  (d[0][0] %.4f ref %.4f; d[77][71] %.4f ref %.4f)\n", worst, bad,
             128 * 80, hd[0], ref[0], hd[77 * 80 + 71], ref[77 * 80 + 71]);
    else
      printf("one-hot k=%3d: d[0][0] %.1f d[5][3] %.1f d[100][63] %.1f d[31][64] %.1f d[127][71] %.1f d[9][72] %.1f (%d off)\n",
             onehot[test - 1], hd[0], hd[5 * 80 + 3], hd[100 * 80 + 63], hd[31 * 80 + 64], hd[127 * 80 + 71], hd[9 * 80 + 72], bad);
    bad_total += bad;
  }
  printf(bad_total == 0 ? "PROBE OK: A-from-TMEM layout as assumed\n" : "PROBE MISMATCH\n");
  return bad_total != 0;
}
