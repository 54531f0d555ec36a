// cta_group::1 instantiations of the tcgen05 GEMM (split from the planner so the two compile in parallel).
#include "gemm_tcgen05.cuh"

namespace reed {

int gemm_tc_launch_cg1(int bn, int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd,
                       int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st, int grid, int stream_k,
                       const EpiMaps* em) {
  return launch_cg<1>(bn, a_mn, b_mn, ma, mb, D, ldd, d_dtype, M, N, K, ep, st, grid, stream_k, em);
}

// grouped operands (GroupMaps): fp32 D, K-major A, cta_group::1 - the adaLN-Zero modulation linears of all blocks at once
int gemm_tc_launch_grouped(int bn, int b_mn, const GroupMaps& gm, void* D, int64_t ldd, int M, int N, int K, const EpiParams& ep,
                           cudaStream_t st, int grid, int stream_k) {
  const CUtensorMap& ma = gm.a[0];
  const CUtensorMap& mb = gm.b[0];
  if (bn == 256) {
    if (b_mn) return launch<1, 256, 0, 1, float, GroupMaps>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, nullptr, &gm);
    return launch<1, 256, 0, 0, float, GroupMaps>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, nullptr, &gm);
  }
  if (bn == 128) {
    if (b_mn) return launch<1, 128, 0, 1, float, GroupMaps>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, nullptr, &gm);
    return launch<1, 128, 0, 0, float, GroupMaps>(ma, mb, D, ldd, M, N, K, ep, st, grid, stream_k, nullptr, &gm);
  }
  return fail("gemm_tcgen05 (grouped): no kernel for BN=%d", bn);
}

}  // namespace reed
