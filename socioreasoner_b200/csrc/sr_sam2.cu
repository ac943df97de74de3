// sr_sam2.cu -- the SAM2 image path's kernels that are not GEMMs or the big attention (those reuse sr_gemm.cu / sr_attention.cu):
// image transform, Hiera patch embedding, LayerNorm, 2x2 max-pool over index tables, FPN top-down add, and the small pieces of the
// prompt encoder / two-way mask decoder (prompt embedding, token <-> image attention, transposed-conv pixel shuffles, hypernetwork
// mask product).  What they replace: ``SAM2ImagePredictor.set_image`` / ``predict`` as driven by
// roll/distributed/strategy/seg_strategy.py:47-58 (the ``sam2`` package; arithmetic pinned on its HF port,
// transformers/models/sam2/modeling_sam2.py).  All are HBM-/latency-bound element-wise or small-reduction kernels; activations bf16,
// statistics fp32.
#include "sr_host.h"

namespace sr {

// ------------------------------------------------------------------------------------------------ image transform
// uint8 HWC -> fp32 CHW [3, S, S]: ToTensor (/255), bilinear resize (align_corners = false), ImageNet normalisation.
__global__ void sam2_preprocess_kernel(const uint8_t* __restrict__ img, int H, int W, int S, float m0, float m1, float m2, float d0, float d1,
                                       float d2, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= S) return;
  const float sy = fmaxf((y + 0.5f) * (static_cast<float>(H) / S) - 0.5f, 0.f);
  const float sx = fmaxf((x + 0.5f) * (static_cast<float>(W) / S) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), H - 1), x0 = min(static_cast<int>(sx), W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float mean[3] = {m0, m1, m2}, sd[3] = {d0, d1, d2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = img[(static_cast<int64_t>(y0) * W + x0) * 3 + c] * (1.f / 255.f), b = img[(static_cast<int64_t>(y0) * W + x1) * 3 + c] * (1.f / 255.f);
    const float e = img[(static_cast<int64_t>(y1) * W + x0) * 3 + c] * (1.f / 255.f), f = img[(static_cast<int64_t>(y1) * W + x1) * 3 + c] * (1.f / 255.f);
    const float v = (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * e + lx * f);
    out[(static_cast<int64_t>(c) * S + y) * S + x] = (v - mean[c]) / sd[c];
  }
}

// ------------------------------------------------------------------------------------------------ patch embedding
// Conv2d(3 -> C, k x k, stride, padding) + bias + positional table, output rows permuted (window-major order): CTA = 8 x 8 output
// pixels (their input patch staged in shared memory), 256 threads = 64 pixels x 4 channel phases.
constexpr int PE_T = 8;
__global__ void __launch_bounds__(256) sam2_patch_embed_kernel(const float* __restrict__ px, int S, int C, int k, int stride, int pad,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               const float* __restrict__ pos, const int* __restrict__ rows,
                                                               __nv_bfloat16* __restrict__ out) {
  extern __shared__ float pe_smem[];
  const int R = S / stride;
  const int tile_in = (PE_T - 1) * stride + k;
  const int oy0 = blockIdx.y * PE_T, ox0 = blockIdx.x * PE_T;
  const int iy0 = oy0 * stride - pad, ix0 = ox0 * stride - pad;
  for (int i = threadIdx.x; i < 3 * tile_in * tile_in; i += blockDim.x) {
    const int c = i / (tile_in * tile_in), r = (i / tile_in) % tile_in, q = i % tile_in;
    const int y = iy0 + r, x = ix0 + q;
    pe_smem[i] = (y >= 0 && y < S && x >= 0 && x < S) ? px[(static_cast<int64_t>(c) * S + y) * S + x] : 0.f;
  }
  __syncthreads();
  const int p = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int py = p / PE_T, pxx = p % PE_T;
  const int oy = oy0 + py, ox = ox0 + pxx;
  if (oy >= R || ox >= R) return;
  const int row = rows[oy * R + ox];
  const int kk = k * k;
  for (int co = ph; co < C; co += 4) {
    const float* wc = w + static_cast<int64_t>(co) * 3 * kk;
    float acc = bias[co];
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < k; ++r) {
        const float* src = pe_smem + (c * tile_in + py * stride + r) * tile_in + pxx * stride;
        const float* wr = wc + c * kk + r * k;
        for (int q = 0; q < k; ++q) acc += src[q] * __ldg(wr + q);
      }
    out[static_cast<int64_t>(row) * C + co] = __float2bfloat16(acc + pos[static_cast<int64_t>(row) * C + co]);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// y = (x - mean) / sqrt(var + eps) * w + b over the last dimension; warp per row, two passes over the (L1-resident) row.
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int M, int C, const float* __restrict__ w,
                                 const float* __restrict__ b, float eps, __nv_bfloat16* __restrict__ out, int64_t ldo) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const __nv_bfloat16* xr = x + static_cast<int64_t>(row) * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __bfloat162float(xr[c]);
  const float mean = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = __bfloat162float(xr[c]) - mean;
    v += d * d;
  }
  const float inv = rsqrtf(warp_sum(v) / C + eps);
  __nv_bfloat16* o = out + static_cast<int64_t>(row) * ldo;
  for (int c = lane; c < C; c += 32) o[c] = __float2bfloat16((__bfloat162float(xr[c]) - mean) * inv * w[c] + b[c]);
}

// ------------------------------------------------------------------------------------------------ 2x2 max-pool via a source table
__global__ void pool4_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const int* __restrict__ idx, int n, int C,
                             __nv_bfloat16* __restrict__ out) {
  const int row = blockIdx.x;
  const int4 s = reinterpret_cast<const int4*>(idx)[row];
  const __nv_bfloat16 *a = x + static_cast<int64_t>(s.x) * ldx, *b = x + static_cast<int64_t>(s.y) * ldx, *c = x + static_cast<int64_t>(s.z) * ldx,
                      *d = x + static_cast<int64_t>(s.w) * ldx;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    const float v = fmaxf(fmaxf(__bfloat162float(a[j]), __bfloat162float(b[j])), fmaxf(__bfloat162float(c[j]), __bfloat162float(d[j])));
    out[static_cast<int64_t>(row) * C + j] = __float2bfloat16(v);
  }
}

// out[(y, x)] = a[(y, x)] + b[(y / 2, x / 2)]   (FPN top-down: nearest 2x up-sampling of the coarser level)
__global__ void upsample2x_add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int r, int C,
                                      __nv_bfloat16* __restrict__ out) {
  const int row = blockIdx.x, R = 2 * r;
  const int y = row / R, x = row % R;
  const __nv_bfloat16* br = b + static_cast<int64_t>((y >> 1) * r + (x >> 1)) * C;
  for (int j = threadIdx.x; j < C; j += blockDim.x)
    out[static_cast<int64_t>(row) * C + j] = __float2bfloat16(__bfloat162float(a[static_cast<int64_t>(row) * C + j]) + __bfloat162float(br[j]));
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int64_t n, __nv_bfloat16* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}
__global__ void relu_kernel(__nv_bfloat16* x, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __float2bfloat16(fmaxf(__bfloat162float(x[i]), 0.f));
}
__global__ void sigmoid_kernel(float* x, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] = 1.f / (1.f + __expf(-x[i]));
}

// ------------------------------------------------------------------------------------------------ prompt embedding
// Sam2PromptEncoder._embed_points / _embed_boxes: random-Fourier positional encoding of (coord + 0.5) / image_size, plus the label's
// embedding (1 / 0: foreground / background point, 2 / 3: box corners); label -1 = the padding "not a point" token.
__global__ void sam2_prompt_embed_kernel(const float* __restrict__ coords, const int* __restrict__ labels, const float* __restrict__ gauss,
                                         const float* __restrict__ point_embed, const float* __restrict__ not_a_point, int half,
                                         float image_size, __nv_bfloat16* __restrict__ out) {
  const int p = blockIdx.x, lab = labels[p];
  const float cx = 2.f * ((coords[2 * p] + 0.5f) / image_size) - 1.f, cy = 2.f * ((coords[2 * p + 1] + 0.5f) / image_size) - 1.f;
  for (int j = threadIdx.x; j < 2 * half; j += blockDim.x) {
    float v;
    if (lab == -1) {
      v = not_a_point[j];
    } else {
      const int f = j < half ? j : j - half;
      const float ang = 6.283185307179586f * (cx * gauss[f] + cy * gauss[half + f]);
      v = (j < half ? sinf(ang) : cosf(ang)) + point_embed[lab * 2 * half + j];
    }
    out[static_cast<int64_t>(p) * 2 * half + j] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------ small attention
// softmax(q k^T / sqrt(d)) v for the mask decoder (head_dim 16 / 32; one side has a handful of rows).  One warp per (query row, head):
// every lane walks the keys lane, lane + 32, ... with its own online softmax, the 32 partial states are merged at the end.
template <int D>
__global__ void small_attention_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
                                       const __nv_bfloat16* __restrict__ v, int64_t ldv, int Mq, int Mk, int H, float scale,
                                       __nv_bfloat16* __restrict__ out, int64_t ldo) {
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= Mq * H) return;
  const int row = wid / H, h = wid % H;
  float qv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) qv[d] = __bfloat162float(q[static_cast<int64_t>(row) * ldq + h * D + d]) * scale;
  float m = -INFINITY, l = 0.f, o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
  for (int j = lane; j < Mk; j += 32) {
    const __nv_bfloat16* kr = k + static_cast<int64_t>(j) * ldk + h * D;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) s += qv[d] * __bfloat162float(kr[d]);
    const float mn = fmaxf(m, s), a = __expf(m - mn), e = __expf(s - mn);
    const __nv_bfloat16* vr = v + static_cast<int64_t>(j) * ldv + h * D;
    l = l * a + e;
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = o[d] * a + e * __bfloat162float(vr[d]);
    m = mn;
  }
  const float M = warp_max(m);
  const float f = (m == -INFINITY) ? 0.f : __expf(m - M);
  const float L = warp_sum(l * f);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float od = warp_sum(o[d] * f);
    if (lane == 0) out[static_cast<int64_t>(row) * ldo + h * D + d] = __float2bfloat16(od / L);
  }
}

// ------------------------------------------------------------------------------------------------ transposed-conv pixel shuffle
// y [r*r, 4*C] = GEMM output of ConvTranspose2d(k = 2, s = 2) with columns (dy, dx, c); out row (2y + dy, 2x + dx):
//   z = y[(y, x)][(dy, dx, :)] + bias + feat[row];  LN: z = LayerNorm_C(z) (channels-first LayerNorm of the reference);  out = gelu(z).
// Warp per output row (C = 32 or 64).
template <bool LN>
__global__ void sam2_upscale_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ feat, const float* __restrict__ bias,
                                    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int r, int C,
                                    __nv_bfloat16* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int R = 2 * r;
  if (row >= R * R) return;
  const int oy = row / R, ox = row % R;
  const __nv_bfloat16* src = y + static_cast<int64_t>((oy >> 1) * r + (ox >> 1)) * 4 * C + ((oy & 1) * 2 + (ox & 1)) * C;
  float z[2];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + 32 * i;
    z[i] = c < C ? __bfloat162float(src[c]) + bias[c] + __bfloat162float(feat[static_cast<int64_t>(row) * C + c]) : 0.f;
    s += z[i];
  }
  if (LN) {
    const float mean = warp_sum(s) / C;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float d = (lane + 32 * i < C) ? z[i] - mean : 0.f;
      v += d * d;
    }
    const float inv = rsqrtf(warp_sum(v) / C + eps);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      if (c < C) z[i] = (z[i] - mean) * inv * ln_w[c] + ln_b[c];
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + 32 * i;
    if (c < C) out[static_cast<int64_t>(row) * C + c] = __float2bfloat16(0.5f * z[i] * (1.f + erff(z[i] * 0.7071067811865476f)));
  }
}

// out[m][p] = sum_c hyper[first + m][c] * up[p][c]   (C = 32: one thread per pixel, hypernetwork rows in shared memory)
__global__ void sam2_mask_logits_kernel(const __nv_bfloat16* __restrict__ hyper, const __nv_bfloat16* __restrict__ up, int first, int count,
                                        int C, int64_t N, float* __restrict__ out) {
  __shared__ float hs[8 * 64];
  for (int i = threadIdx.x; i < count * C; i += blockDim.x) hs[i] = __bfloat162float(hyper[(first + i / C) * C + i % C]);
  __syncthreads();
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= N) return;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  for (int c = 0; c < C; ++c) {
    const float u = __bfloat162float(up[p * C + c]);
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (m < count) acc[m] += hs[m * C + c] * u;
  }
#pragma unroll
  for (int m = 0; m < 8; ++m)
    if (m < count) out[static_cast<int64_t>(m) * N + p] = acc[m];
}

}  // namespace sr

using namespace sr;
#define SAM_STREAM cudaStream_t s = reinterpret_cast<cudaStream_t>(stream)

extern "C" int sr_sam2_preprocess(const uint8_t* image, int height, int width, int size, const float* mean3, const float* std3, float* out,
                                  void* stream) {
  SR_REQUIRE(image && out && mean3 && std3 && height > 0 && width > 0 && size > 0, "sr_sam2_preprocess: bad arguments");
  SAM_STREAM;
  sam2_preprocess_kernel<<<dim3((size + 127) / 128, size), 128, 0, s>>>(image, height, width, size, mean3[0], mean3[1], mean3[2], std3[0],
                                                                        std3[1], std3[2], out);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_sam2_patch_embed(const float* pixels, int size, int channels, int kernel, int stride, int padding, const float* weight,
                                   const float* bias, const float* pos_table, const int32_t* dst_rows, void* out, void* stream) {
  SR_REQUIRE(pixels && weight && bias && pos_table && dst_rows && out, "sr_sam2_patch_embed: null pointer");
  SR_REQUIRE(size % stride == 0 && kernel > 0 && kernel <= 16, "sr_sam2_patch_embed: bad geometry");
  SAM_STREAM;
  const int R = size / stride, tile_in = (PE_T - 1) * stride + kernel;
  const size_t smem = 3 * tile_in * tile_in * sizeof(float);
  sam2_patch_embed_kernel<<<dim3((R + PE_T - 1) / PE_T, (R + PE_T - 1) / PE_T), 256, smem, s>>>(
      pixels, size, channels, kernel, stride, padding, weight, bias, pos_table, dst_rows, reinterpret_cast<__nv_bfloat16*>(out));
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_layernorm_bf16(const void* x, int64_t ldx, int rows, int cols, const float* weight, const float* bias, float eps, void* out,
                                 int64_t ldo, void* stream) {
  SR_REQUIRE(x && weight && bias && out && rows > 0 && cols > 0, "sr_layernorm_bf16: bad arguments");
  SAM_STREAM;
  layernorm_kernel<<<(rows + 7) / 8, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, cols, weight, bias, eps,
                                                  reinterpret_cast<__nv_bfloat16*>(out), ldo);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_pool4_rows(const void* x, int64_t ldx, const int32_t* idx4, int n_out, int cols, void* out, void* stream) {
  SR_REQUIRE(x && idx4 && out && n_out > 0 && cols > 0 && (reinterpret_cast<uintptr_t>(idx4) & 15) == 0, "sr_pool4_rows: bad arguments");
  SAM_STREAM;
  pool4_kernel<<<n_out, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, idx4, n_out, cols, reinterpret_cast<__nv_bfloat16*>(out));
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_upsample2x_add(const void* fine, const void* coarse, int coarse_res, int cols, void* out, void* stream) {
  SR_REQUIRE(fine && coarse && out && coarse_res > 0 && cols > 0, "sr_upsample2x_add: bad arguments");
  SAM_STREAM;
  upsample2x_add_kernel<<<4 * coarse_res * coarse_res, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(fine),
                                                                    reinterpret_cast<const __nv_bfloat16*>(coarse), coarse_res, cols,
                                                                    reinterpret_cast<__nv_bfloat16*>(out));
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

/* op: 0 = out = a + b (bf16), 1 = relu in place on a (bf16), 2 = sigmoid in place on a (fp32) */
extern "C" int sr_elementwise(int op, void* a, const void* b, void* out, int64_t n, void* stream) {
  SR_REQUIRE(a && n > 0 && op >= 0 && op <= 2 && (op != 0 || (b && out)), "sr_elementwise: bad arguments");
  SAM_STREAM;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  if (op == 0)
    add_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), n,
                                    reinterpret_cast<__nv_bfloat16*>(out));
  else if (op == 1)
    relu_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(a), n);
  else
    sigmoid_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<float*>(a), n);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_sam2_prompt_embed(const float* coords, const int32_t* labels, int n_points, const float* gauss, const float* point_embed,
                                    const float* not_a_point, int hidden, float image_size, void* out, void* stream) {
  SR_REQUIRE(coords && labels && gauss && point_embed && not_a_point && out && n_points > 0 && hidden % 2 == 0, "sr_sam2_prompt_embed: bad arguments");
  SAM_STREAM;
  sam2_prompt_embed_kernel<<<n_points, 128, 0, s>>>(coords, labels, gauss, point_embed, not_a_point, hidden / 2, image_size,
                                                    reinterpret_cast<__nv_bfloat16*>(out));
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_small_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int n_q, int n_k,
                                  int heads, int head_dim, void* out, int64_t ldo, void* stream) {
  SR_REQUIRE(q && k && v && out && n_q > 0 && n_k > 0 && heads > 0, "sr_small_attention: bad arguments");
  SR_REQUIRE(head_dim == 16 || head_dim == 32, "sr_small_attention: head_dim must be 16 or 32 (got %d)", head_dim);
  SAM_STREAM;
  const int warps = n_q * heads;
  const float scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  auto Q = reinterpret_cast<const __nv_bfloat16*>(q);
  auto K = reinterpret_cast<const __nv_bfloat16*>(k);
  auto V = reinterpret_cast<const __nv_bfloat16*>(v);
  auto O = reinterpret_cast<__nv_bfloat16*>(out);
  if (head_dim == 16)
    small_attention_kernel<16><<<(warps + 3) / 4, 128, 0, s>>>(Q, ldq, K, ldk, V, ldv, n_q, n_k, heads, scale, O, ldo);
  else
    small_attention_kernel<32><<<(warps + 3) / 4, 128, 0, s>>>(Q, ldq, K, ldk, V, ldv, n_q, n_k, heads, scale, O, ldo);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_sam2_upscale(const void* y, const void* feat, const float* bias, const float* ln_weight, const float* ln_bias, float eps,
                               int in_res, int channels, void* out, void* stream) {
  SR_REQUIRE(y && feat && bias && out && in_res > 0 && channels > 0 && channels <= 64, "sr_sam2_upscale: bad arguments (channels <= 64)");
  SAM_STREAM;
  const int rows = 4 * in_res * in_res;
  auto Y = reinterpret_cast<const __nv_bfloat16*>(y);
  auto F = reinterpret_cast<const __nv_bfloat16*>(feat);
  auto O = reinterpret_cast<__nv_bfloat16*>(out);
  if (ln_weight != nullptr)
    sam2_upscale_kernel<true><<<(rows + 7) / 8, 256, 0, s>>>(Y, F, bias, ln_weight, ln_bias, eps, in_res, channels, O);
  else
    sam2_upscale_kernel<false><<<(rows + 7) / 8, 256, 0, s>>>(Y, F, bias, nullptr, nullptr, eps, in_res, channels, O);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_sam2_mask_logits(const void* hyper, const void* upscaled, int first, int count, int channels, int64_t n_pixels, float* out,
                                   void* stream) {
  SR_REQUIRE(hyper && upscaled && out && count > 0 && count <= 8 && channels > 0 && channels <= 64 && n_pixels > 0, "sr_sam2_mask_logits: bad arguments");
  SAM_STREAM;
  sam2_mask_logits_kernel<<<static_cast<unsigned>((n_pixels + 255) / 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(hyper),
                                                                                      reinterpret_cast<const __nv_bfloat16*>(upscaled), first,
                                                                                      count, channels, n_pixels, out);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}
