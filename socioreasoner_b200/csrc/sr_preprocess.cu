// sr_preprocess.cu -- uint8 tile -> Qwen2.5-VL pixel_values rows on the device.
//
// What it replaces: the CPU image path in front of the ViT -- PIL's 8-bit bicubic resize on the driver
// (``process_image``, roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py:126-144) followed by the HF processor the collator
// calls (roll/datasets/collator.py:444-565 -> Qwen2VLImageProcessor: rescale 1/255, normalise, duplicate the frame to T = 2,
// patchify into 2x2-merge-block row order, columns (C, T, Ph, Pw)).  With it a rank uploads 3 MB of uint8 per 1024^2 tile instead
// of 13.7 MB of fp32 pixel rows.  Results are BIT-EXACT with that CPU path:
//   * PIL's resampler (libImaging/Resample.c) is integer arithmetic: two separable passes, horizontal then vertical, each
//     out = clip8((2^21 + sum_k in[first + k] * coeff[k]) >> 22) on uint8 data with 22-bit fixed-point coefficient tables that
//     depend only on (in_size, out_size) -- the tables are built on the host (socioreasoner_b200/preprocess.py) and passed in;
//   * the fused rescale + normalise is (float(u8) - mean_s) / std_s with IEEE fp32 subtract and divide.
// HBM-bound (3 B read, 2 x 1176 x 4 / 196 ~ 48 B written per output pixel); a tile is ~20 MB of traffic.
#include "sr_host.h"

namespace sr {

constexpr int PP_PRECISION_BITS = 22;

SR_DEVINL int pp_clip8(int v) {
  v >>= PP_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: in [H, W_in, 3] -> out [H, W_out, 3]; thread = one output pixel, all 3 channels
__global__ void pp_resize_h_kernel(const uint8_t* __restrict__ in, int H, int W_in, int W_out, const int* __restrict__ bounds,
                                   const int* __restrict__ coeffs, int ksize, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W_out) return;
  const int xmin = bounds[2 * x], n = bounds[2 * x + 1];
  const int* k = coeffs + static_cast<int64_t>(x) * ksize;
  const uint8_t* src = in + (static_cast<int64_t>(y) * W_in + xmin) * 3;
  int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int j = 0; j < n; ++j) {
    const int c = k[j];
    s0 += src[3 * j] * c;
    s1 += src[3 * j + 1] * c;
    s2 += src[3 * j + 2] * c;
  }
  uint8_t* d = out + (static_cast<int64_t>(y) * W_out + x) * 3;
  d[0] = static_cast<uint8_t>(pp_clip8(s0));
  d[1] = static_cast<uint8_t>(pp_clip8(s1));
  d[2] = static_cast<uint8_t>(pp_clip8(s2));
}

// vertical pass (optional) + rescale / normalise + T = 2 duplicate + merge-block patchify.
// in [H_in, W, 3] -> rows of out: patch (gy, gx) -> row ((gy/2) * (gw/2) + gx/2) * 4 + (gy%2) * 2 + gx%2 (then row_map, if given);
// column c * (2 * P * P) + t * P * P + py * P + px.  thread = one output pixel.
template <typename OutT>
__global__ void pp_patchify_kernel(const uint8_t* __restrict__ in, int H_in, int W, int H_out, const int* __restrict__ bounds,
                                   const int* __restrict__ coeffs, int ksize, float m0, float m1, float m2, float d0, float d1, float d2,
                                   OutT* __restrict__ out, int64_t ld_out, const int* __restrict__ row_map, int patch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  int v0, v1, v2;
  if (bounds != nullptr) {
    const int ymin = bounds[2 * y], n = bounds[2 * y + 1];
    const int* k = coeffs + static_cast<int64_t>(y) * ksize;
    int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < n; ++j) {
      const uint8_t* src = in + (static_cast<int64_t>(ymin + j) * W + x) * 3;
      const int c = k[j];
      s0 += src[0] * c;
      s1 += src[1] * c;
      s2 += src[2] * c;
    }
    v0 = pp_clip8(s0); v1 = pp_clip8(s1); v2 = pp_clip8(s2);
  } else {
    const uint8_t* src = in + (static_cast<int64_t>(y) * W + x) * 3;
    v0 = src[0]; v1 = src[1]; v2 = src[2];
  }
  const float f0 = __fdiv_rn(__fsub_rn(static_cast<float>(v0), m0), d0);
  const float f1 = __fdiv_rn(__fsub_rn(static_cast<float>(v1), m1), d1);
  const float f2 = __fdiv_rn(__fsub_rn(static_cast<float>(v2), m2), d2);
  const int gy = y / patch, py = y - gy * patch, gx = x / patch, px = x - gx * patch;
  const int gw = W / patch;
  int row = ((gy >> 1) * (gw >> 1) + (gx >> 1)) * 4 + (gy & 1) * 2 + (gx & 1);
  if (row_map != nullptr) row = row_map[row];
  const int pp = patch * patch;
  OutT* o = out + static_cast<int64_t>(row) * ld_out + py * patch + px;
  o[0] = static_cast<OutT>(f0); o[pp] = static_cast<OutT>(f0);
  o[2 * pp] = static_cast<OutT>(f1); o[3 * pp] = static_cast<OutT>(f1);
  o[4 * pp] = static_cast<OutT>(f2); o[5 * pp] = static_cast<OutT>(f2);
}

}  // namespace sr

using namespace sr;

extern "C" int sr_preprocess_tile(const uint8_t* image, int height, int width, int out_height, int out_width,
                                  const int32_t* h_bounds, const int32_t* h_coeffs, int h_ksize, const int32_t* v_bounds,
                                  const int32_t* v_coeffs, int v_ksize, uint8_t* scratch, const float* mean3, const float* std3,
                                  void* pixel_values, int out_dtype, int64_t ld_out, const int32_t* row_map, int patch_size,
                                  void* stream) {
  SR_REQUIRE(image && pixel_values && mean3 && std3, "sr_preprocess_tile: null pointer");
  SR_REQUIRE(height > 0 && width > 0 && out_height > 0 && out_width > 0 && patch_size > 0, "sr_preprocess_tile: sizes");
  SR_REQUIRE(out_height % (2 * patch_size) == 0 && out_width % (2 * patch_size) == 0,
             "sr_preprocess_tile: output size %dx%d is not a multiple of 2 * patch_size", out_height, out_width);
  SR_REQUIRE(ld_out >= 6LL * patch_size * patch_size, "sr_preprocess_tile: ld_out");
  SR_REQUIRE(out_dtype == SR_DTYPE_F32 || out_dtype == SR_DTYPE_BF16, "sr_preprocess_tile: out_dtype");
  const bool need_h = out_width != width, need_v = out_height != height;
  SR_REQUIRE(!need_h || (h_bounds && h_coeffs && h_ksize > 0 && scratch), "sr_preprocess_tile: horizontal tables / scratch missing");
  SR_REQUIRE(!need_v || (v_bounds && v_coeffs && v_ksize > 0), "sr_preprocess_tile: vertical tables missing");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint8_t* src = image;
  if (need_h) {
    dim3 grid((out_width + 127) / 128, height);
    pp_resize_h_kernel<<<grid, 128, 0, s>>>(image, height, width, out_width, h_bounds, h_coeffs, h_ksize, scratch);
    SR_CHECK_CUDA(cudaGetLastError());
    src = scratch;
  }
  dim3 grid((out_width + 127) / 128, out_height);
  const float h[6] = {mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]};  // HOST arrays (kernel arguments)
  if (out_dtype == SR_DTYPE_F32) {
    pp_patchify_kernel<float><<<grid, 128, 0, s>>>(src, height, out_width, out_height, need_v ? v_bounds : nullptr, v_coeffs, v_ksize, h[0],
                                                   h[1], h[2], h[3], h[4], h[5], reinterpret_cast<float*>(pixel_values), ld_out, row_map,
                                                   patch_size);
  } else {
    pp_patchify_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(src, height, out_width, out_height, need_v ? v_bounds : nullptr, v_coeffs,
                                                           v_ksize, h[0], h[1], h[2], h[3], h[4], h[5],
                                                           reinterpret_cast<__nv_bfloat16*>(pixel_values), ld_out, row_map, patch_size);
  }
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}
