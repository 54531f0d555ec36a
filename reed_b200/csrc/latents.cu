// Latent data path of the train step: the VAE-posterior draw that turns a batch of stored
// SD-VAE moments into the latents SILoss consumes.
//
// Reference: /root/reference/image/train.py:84-91 (sample_posterior): mean, std = chunk(moments, 2, dim=1);
// z = mean + std * randn_like(mean); z = z * latents_scale + latents_bias, with latents_scale / latents_bias the
// per-channel [1,C,1,1] tensors of train.py:226-231.  PyTorch runs it as chunk views + 4 elementwise kernels
// (5 reads + 4 writes of a latent-sized tensor); here it is one pass: read mean, std, noise, write z.
//
// Every product and sum is rounded separately (__fmul_rn / __fadd_rn, no FMA contraction) so the result is
// bit-identical to the reference's sequence of PyTorch fp32 kernels for the same noise tensor.
#include "common.cuh"

namespace reed {

struct PosteriorArgs {
  const float* moments;   // [B, 2C, HW]: channels [0,C) = mean, [C,2C) = std
  const float* noise;     // [B, C, HW] standard normals
  const float* scale;     // [C] or null (then scale_s)
  const float* bias;      // [C] or null (then bias_s)
  float* out;             // [B, C, HW]
  int64_t total;          // B*C*HW
  int chw, hw;
  float scale_s, bias_s;
};

__device__ __forceinline__ float posterior_one(float mean, float sd, float n, float sc, float bi) {
  return __fadd_rn(__fmul_rn(__fadd_rn(mean, __fmul_rn(sd, n)), sc), bi);
}

// VEC = 4: HW % 4 == 0, so a group of 4 elements never straddles a channel and every access is 16 bytes
template <int VEC>
__global__ void __launch_bounds__(256) sample_posterior_kernel(PosteriorArgs a) {
  const int64_t groups = a.total / VEC;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = g * VEC;
    const int64_t b = i / a.chw;
    const int r = (int)(i - b * a.chw);
    const int c = r / a.hw;
    const float sc = a.scale ? a.scale[c] : a.scale_s;
    const float bi = a.bias ? a.bias[c] : a.bias_s;
    const float* mean = a.moments + b * 2 * a.chw + r;
    const float* sd = mean + a.chw;
    if (VEC == 4) {
      F4 m = load4(mean), s = load4(sd), n = load4(a.noise + i), o;
#pragma unroll
      for (int k = 0; k < 4; ++k) o.v[k] = posterior_one(m.v[k], s.v[k], n.v[k], sc, bi);
      store4(a.out + i, o);
    } else {
      a.out[i] = posterior_one(*mean, *sd, a.noise[i], sc, bi);
    }
  }
}

}  // namespace reed

using namespace reed;

extern "C" int reed_sample_posterior(const void* moments, const void* noise, const void* scale, const void* bias,
                                     float scale_scalar, float bias_scalar, void* out, int batch, int channels, int hw,
                                     void* stream) {
  REED_REQUIRE(batch >= 0 && channels > 0 && hw > 0, "sample_posterior: bad shape");
  PosteriorArgs a{(const float*)moments, (const float*)noise, (const float*)scale, (const float*)bias, (float*)out,
                  (int64_t)batch * channels * hw, channels * hw, hw, scale_scalar, bias_scalar};
  if (a.total == 0) return 0;
  const bool vec = hw % 4 == 0 && ((uintptr_t)moments | (uintptr_t)noise | (uintptr_t)out) % 16 == 0;
  const int64_t work = vec ? a.total / 4 : a.total;
  int64_t blocks = (work + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (vec) sample_posterior_kernel<4><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  else sample_posterior_kernel<1><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  REED_LAUNCH_CHECK();
  return 0;
}
