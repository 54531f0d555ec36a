// Probe (round 2): how fast can one SM pull the attention operand tiles through TMA?  Every CTA streams [128 x hd] tiles of
// the packed qkv tensor [B*T, 3H, hd] (the access pattern of the attention kernels: 128 rows of 144 bytes, 6912 bytes apart)
// into a ring of shared-memory stages and does nothing else; reports cycles per tile for head_dim 64 (one SWIZZLE_128B box
// per tile) and 72 (main box + 16-column SWIZZLE_32B tail box), with 1 / 4 / 8 stages in flight, plus the TMA-store rate.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I reed_b200/csrc profiles/probe_tma_rate.cu -o profiles/bin/probe_tma -lcuda
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "attention_fa.cuh"

namespace reed {
thread_local char g_err[512];
int fail(const char* fmt, ...) { fprintf(stderr, "fail: %s\n", fmt); return 1; }
}  // namespace reed
using namespace reed;
using namespace reed::fa;

template <int HD, int STAGES, bool STORE>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ AttnMaps maps, int T, int H, int tiles_per_cta,
                                                       long long* cycles) {
  using TL = Tile<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * TL::kBytes);
  if (threadIdx.x == 0) {
    for (int k = 0; k < STAGES; ++k) mbar_init(bars + k, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const int nq = T / kRows;
    for (int i = 0; i < tiles_per_cta + STAGES; ++i) {
      const int st = i % STAGES;
      if (i >= STAGES) {
        mbar_wait(bars + st, (uint32_t)((i - STAGES) / STAGES) & 1u);      // tile i - STAGES has landed
        if (STORE) {
          const int j = (i - STAGES) * gridDim.x + blockIdx.x;
          fence_proxy_async();
          store_tile<HD>(&maps.out_main, &maps.out_tail8, sbase + st * TL::kBytes, j % (3 * H), (j / (3 * H)) * kRows);
          tma_store_commit();
          tma_store_wait_read();
        }
      }
      if (i < tiles_per_cta) {
        const int j = i * gridDim.x + blockIdx.x;       // tile index: (row block, head column)
        const int hcol = j % (3 * H), rb = j / (3 * H);
        mbar_expect_tx(bars + st, TL::kBytes);
        load_tile<HD>(&maps.qkv_main, &maps.qkv_tail, bars + st, sbase + st * TL::kBytes, hcol, rb * kRows);
      }
    }
    if (STORE) tma_store_wait_all();
    cycles[blockIdx.x] = clock64() - t0;
    (void)nq;
  }
}

template <int HD, int STAGES, bool STORE>
void run(const char* label, int B, int T, int H) {
  const int64_t rows = (int64_t)B * T;
  const size_t bytes = (size_t)rows * 3 * H * HD * 2;
  void *src, *dst;
  cudaMalloc(&src, bytes);
  cudaMalloc(&dst, bytes);
  cudaMemset(src, 1, bytes);
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  make_map3(&maps.qkv_main, src, rows, 3 * H, HD, 0);
  make_map3(&maps.out_main, dst, rows, 3 * H, HD, 0);
  if (Tile<HD>::kTail) {
    make_map3(&maps.qkv_tail, src, rows, 3 * H, HD, 1);
    make_map3(&maps.out_tail8, dst, rows, 3 * H, HD, 2);
  }
  const int total_tiles = (int)(rows / kRows) * 3 * H;
  const int grid = 148, per = total_tiles / grid;
  long long* cyc;
  cudaMalloc(&cyc, grid * sizeof(long long));
  const int smem = 1024 + STAGES * Tile<HD>::kBytes + 256;
  cudaFuncSetAttribute(stream_kernel<HD, STAGES, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // flush L2
  void* flush;
  cudaMalloc(&flush, 256u << 20);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(flush, rep, 256u << 20);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    stream_kernel<HD, STAGES, STORE><<<grid, 64, smem>>>(maps, T, H, per, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("%s: CUDA error %s\n", label, cudaGetErrorString(err));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    if (rep == 1)
      printf("%-34s hd=%d stages=%d: %5d tiles/CTA, %7.1f us, %6.0f cycles/tile/SM, %6.2f TB/s useful\n", label, HD, STAGES, per,
             ms * 1e3, (double)mx / per, (double)per * grid * kRows * HD * 2 * (STORE ? 2 : 1) / (ms * 1e-3) / 1e12);
  }
  cudaFree(src);
  cudaFree(dst);
  cudaFree(cyc);
  cudaFree(flush);
}

int main() {
  const int B = 32, T = 256, H = 16;
  run<64, 1, false>("load, cold L2", B, T, H);
  run<64, 4, false>("load, cold L2", B, T, H);
  run<64, 8, false>("load, cold L2", B, T, H);
  run<72, 1, false>("load, cold L2", B, T, H);
  run<72, 4, false>("load, cold L2", B, T, H);
  run<72, 8, false>("load, cold L2", B, T, H);
  run<64, 4, true>("load + store (wait_read each)", B, T, H);
  run<72, 4, true>("load + store (wait_read each)", B, T, H);
  return 0;
}
