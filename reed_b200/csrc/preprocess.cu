// Raw-image preprocessing in front of the frozen target encoders (first half): one pass from the
// uint8 pixels of the data loader to the normalised, bicubic-resized encoder input.
//
// Reference: /root/reference/image/train.py:53-74 (preprocess_raw_image): x / 255, torchvision Normalize
// ((x - mean) / std per channel), F.interpolate(x, 224 * (resolution // 256), mode='bicubic') - normalise-then-resize
// for dinov2 / jepa, resize-then-normalise for clip, normalise only for mocov3 / mae / dinov1.  PyTorch runs it as
// 3-4 kernels over fp32 images (>= 36 B per input pixel); here it is 1 B read per input pixel (from L1/L2 for the 16 taps)
// and 4 B (2 B) written per output pixel.
//
// Bicubic = ATen upsample_bicubic2d: align_corners=False, A = -0.75, source index scale * (dst + 0.5) - 0.5 (not
// clamped), the 4x4 taps clamped to the image, rows interpolated first, then the column.  fp32 arithmetic; parity bar
// 1e-5 absolute against the reference function (values are O(1)).
// NOT YET RUN ON HARDWARE: written after round 1's GPU minutes were spent (tests/test_zz_next_gpu.py).
#include "common.cuh"

namespace reed {

struct PreArgs {
  const void* src;      // [B, C, S, S] uint8 or fp32, values 0..255
  void* dst;            // [B, C, O, O] fp32 or bf16
  int C, S, O;
  int64_t total;        // B * C * O * O
  float scale;          // S / O
  float mean[4], stdv[4];
  int resize_first;     // 1: clip order (resize x/255, then normalise), 0: normalise the taps, then resize
};

__device__ __forceinline__ void cubic_coeffs(float t, float* w) {
  const float A = -0.75f;
  float x = t + 1.f;
  w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

template <typename TS> __device__ __forceinline__ float pixel(const TS* p) { return (float)*p; }

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) preprocess_kernel(PreArgs a) {
  const TS* src = reinterpret_cast<const TS*>(a.src);
  TD* dst = reinterpret_cast<TD*>(a.dst);
  const int64_t plane_out = (int64_t)a.O * a.O, plane_in = (int64_t)a.S * a.S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bc = i / plane_out;
    const int r = (int)(i - bc * plane_out);
    const int oy = r / a.O, ox = r - oy * a.O;
    const int c = (int)(bc % a.C);
    // selects, not a.mean[c]: dynamic indexing would copy the parameter arrays to local memory
    const float mean = c == 0 ? a.mean[0] : (c == 1 ? a.mean[1] : (c == 2 ? a.mean[2] : a.mean[3]));
    const float sd = c == 0 ? a.stdv[0] : (c == 1 ? a.stdv[1] : (c == 2 ? a.stdv[2] : a.stdv[3]));
    const TS* img = src + bc * plane_in;
    float out;
    if (a.O == a.S) {                       // no resize: normalise only
      out = (pixel(img + (int64_t)oy * a.S + ox) / 255.f - mean) / sd;
    } else {
      const float sy = a.scale * ((float)oy + 0.5f) - 0.5f, sx = a.scale * ((float)ox + 0.5f) - 0.5f;
      const float fy = floorf(sy), fx = floorf(sx);
      float wy[4], wx[4];
      cubic_coeffs(sy - fy, wy);
      cubic_coeffs(sx - fx, wx);
      const int iy = (int)fy, ix = (int)fx;
      int xs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xs[k] = min(max(ix - 1 + k, 0), a.S - 1);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const TS* row = img + (int64_t)min(max(iy - 1 + j, 0), a.S - 1) * a.S;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[k] = pixel(row + xs[k]) / 255.f;
          if (!a.resize_first) v[k] = (v[k] - mean) / sd;
        }
        acc += (v[0] * wx[0] + v[1] * wx[1] + v[2] * wx[2] + v[3] * wx[3]) * wy[j];
      }
      out = a.resize_first ? (acc - mean) / sd : acc;
    }
    dst[i] = from_f<TD>(out);
  }
}

}  // namespace reed

using namespace reed;

// src_dtype: 0 = fp32, 2 = uint8.  dst_dtype: 0 = fp32, 1 = bf16.  mean / stdv: HOST arrays of `channels` floats.
extern "C" int reed_preprocess_image(const void* src, int src_dtype, void* dst, int dst_dtype, int batch, int channels,
                                     int in_size, int out_size, const float* mean, const float* stdv, int resize_first,
                                     void* stream) {
  REED_REQUIRE(batch >= 0 && channels >= 1 && channels <= 4 && in_size >= 1 && out_size >= 1, "preprocess_image: bad shape");
  REED_REQUIRE(src_dtype == 0 || src_dtype == 2, "preprocess_image: source must be fp32 (0) or uint8 (2)");
  REED_REQUIRE(dst_dtype == kF32 || dst_dtype == kBF16, "preprocess_image: destination must be fp32 or bf16");
  REED_REQUIRE(mean != nullptr && stdv != nullptr, "preprocess_image: mean / std missing");
  PreArgs a;
  a.src = src; a.dst = dst; a.C = channels; a.S = in_size; a.O = out_size;
  a.total = (int64_t)batch * channels * out_size * out_size;
  a.scale = (float)in_size / (float)out_size;
  for (int c = 0; c < 4; ++c) {
    a.mean[c] = c < channels ? mean[c] : 0.f;
    a.stdv[c] = c < channels ? stdv[c] : 1.f;
  }
  a.resize_first = resize_first;
  if (a.total == 0) return 0;
  int64_t blocks = (a.total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == 2) {
    if (dst_dtype == kBF16) preprocess_kernel<uint8_t, bf16><<<(int)blocks, 256, 0, st>>>(a);
    else preprocess_kernel<uint8_t, float><<<(int)blocks, 256, 0, st>>>(a);
  } else {
    if (dst_dtype == kBF16) preprocess_kernel<float, bf16><<<(int)blocks, 256, 0, st>>>(a);
    else preprocess_kernel<float, float><<<(int)blocks, 256, 0, st>>>(a);
  }
  REED_LAUNCH_CHECK();
  return 0;
}
