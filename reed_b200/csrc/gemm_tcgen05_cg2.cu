// cta_group::2 instantiations of the tcgen05 GEMM (split from the planner so the two compile in parallel).
#include "gemm_tcgen05.cuh"

namespace reed {

int gemm_tc_launch_cg2(int bn, int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, int64_t ldd,
                       int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st, int grid, int stream_k,
                       const EpiMaps* em) {
  return launch_cg<2>(bn, a_mn, b_mn, ma, mb, D, ldd, d_dtype, M, N, K, ep, st, grid, stream_k, em);
}

}  // namespace reed
