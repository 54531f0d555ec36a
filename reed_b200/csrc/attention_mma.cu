// Tensor-core flash attention (bf16).  Placeholder translation unit: the kernel lands in a follow-up commit;
// until then attn_mma_supported() reports false and reed_attn_* run the fp32-math SIMT kernels.
#include "common.cuh"

namespace reed {

bool attn_mma_supported(int T, int hd) { (void)T; (void)hd; return false; }

int attn_mma_fwd(const void*, void*, float*, int, int, int, int, cudaStream_t) {
  return fail("tensor-core attention is not built in this revision");
}
int attn_mma_bwd(const void*, const void*, const void*, const float*, void*, float*, int, int, int, int, cudaStream_t) {
  return fail("tensor-core attention is not built in this revision");
}

}  // namespace reed
