// sr_elementwise.cu -- HBM-bound kernels of the hot path: RMSNorm, rotary tables + apply (ViT 2-D RoPE and
// decoder M-RoPE) fused with the Q/K/V re-layout the attention kernel wants, row gathers (window re-order,
// token embedding + vision scatter).  All coalesced, 16-byte vectorised, fp32 math, bf16 storage.
// Reference call sites are cited per kernel; `HF:` = transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py.
#include "sr_host.h"

namespace sr {

// ------------------------------------------------------------------------------------------------
// RMSNorm (HF:57-71; mcore copy mcore_adapter/src/mcore_adapter/models/model_utils.py:91-106).
// y = w * (x * rsqrt(mean(x^2) + eps)), fp32 math, one warp per row, row kept in registers.
// ------------------------------------------------------------------------------------------------
template <int CHUNKS>  // D = CHUNKS * 256
__global__ void __launch_bounds__(256) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                     const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
                                                     int64_t ldy, int rows, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4 v[CHUNKS];
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) v[i] = __ldg(src + lane + 32 * i);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
      ss += a * a + b * b;
    }
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / static_cast<float>(CHUNKS * 256) + eps);
  uint4* dst = reinterpret_cast<uint4*>(y + static_cast<int64_t>(row) * ldy);
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const int c = (lane + 32 * i) * 8;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
    uint4 o;
    o.x = pack_bf16x2(bf16_lo(v[i].x) * inv * w0.x, bf16_hi(v[i].x) * inv * w0.y);
    o.y = pack_bf16x2(bf16_lo(v[i].y) * inv * w0.z, bf16_hi(v[i].y) * inv * w0.w);
    o.z = pack_bf16x2(bf16_lo(v[i].z) * inv * w1.x, bf16_hi(v[i].z) * inv * w1.y);
    o.w = pack_bf16x2(bf16_lo(v[i].w) * inv * w1.z, bf16_hi(v[i].w) * inv * w1.w);
    dst[lane + 32 * i] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Row gathers.
//  * gather_rows_bf16: dst[i,:] = src[idx[i],:]           (window re-order / its inverse, HF:478-485,511-513;
//                                                           last-token rows for the LM head)
//  * gather_rows_f32_to_bf16: same with fp32 source        (pixel_values arrive fp32 from the HF processor; the
//                                                           window permutation is applied to the patch rows BEFORE the
//                                                           patch-embed GEMM -- it commutes with every per-row op)
//  * embed_scatter: out[i,:] = vis_idx[i] >= 0 ? vis[vis_idx[i],:] : table[ids[i],:]
//                                                          (embed_tokens + masked_scatter, HF:1299-1307)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, int64_t lds,
                                                              const int* __restrict__ idx,
                                                              __nv_bfloat16* __restrict__ dst, int64_t ldd, int rows,
                                                              int vecs /* row length / 8 */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int s = idx ? __ldg(idx + row) : row;
  const uint4* a = reinterpret_cast<const uint4*>(src + static_cast<int64_t>(s) * lds);
  uint4* b = reinterpret_cast<uint4*>(dst + static_cast<int64_t>(row) * ldd);
  for (int i = lane; i < vecs; i += 32) b[i] = __ldg(a + i);
}

__global__ void __launch_bounds__(256) gather_rows_f32_kernel(const float* __restrict__ src, int64_t lds,
                                                             const int* __restrict__ idx,
                                                             __nv_bfloat16* __restrict__ dst, int64_t ldd, int rows,
                                                             int vecs /* row length / 8 */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int s = idx ? __ldg(idx + row) : row;
  const float4* a = reinterpret_cast<const float4*>(src + static_cast<int64_t>(s) * lds);
  uint4* b = reinterpret_cast<uint4*>(dst + static_cast<int64_t>(row) * ldd);
  for (int i = lane; i < vecs; i += 32) {
    const float4 f0 = __ldg(a + 2 * i), f1 = __ldg(a + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(f0.x, f0.y);
    o.y = pack_bf16x2(f0.z, f0.w);
    o.z = pack_bf16x2(f1.x, f1.y);
    o.w = pack_bf16x2(f1.z, f1.w);
    b[i] = o;
  }
}

__global__ void __launch_bounds__(256) embed_scatter_kernel(const __nv_bfloat16* __restrict__ table, int64_t ldt,
                                                           const int64_t* __restrict__ ids,
                                                           const __nv_bfloat16* __restrict__ vis, int64_t ldv,
                                                           const int* __restrict__ vis_idx,
                                                           __nv_bfloat16* __restrict__ out, int64_t ldo, int rows,
                                                           int vecs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int vi = vis_idx ? __ldg(vis_idx + row) : -1;
  const uint4* a = (vi >= 0) ? reinterpret_cast<const uint4*>(vis + static_cast<int64_t>(vi) * ldv)
                             : reinterpret_cast<const uint4*>(table + __ldg(ids + row) * ldt);
  uint4* b = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo);
  for (int i = lane; i < vecs; i += 32) b[i] = __ldg(a + i);
}

// ------------------------------------------------------------------------------------------------
// Rotary tables.
//  * ViT 2-D RoPE (HF:117-130,382-409): freqs = outer(arange, 1/10000^(2i/40)), i < 20; per patch the (h, w)
//    indices pick 20 + 20 values; cos/sin [N,40] fp32 (the upper 40 of HF's 80 are a duplicate).
//  * Decoder M-RoPE (reference mcore_adapter/.../modeling_qwen2_5_vl.py:44-75 + :21-41 section select; HF:595-608,
//    627-669): freq j < 64 uses axis t for j<16, h for 16<=j<40, w for 40<=j<64 (mrope_section [16,24,24]);
//    cos/sin [S,64] fp32 (dims 64..127 duplicate 0..63).
// inv_freq is passed in (computed on the host with the reference's exact fp32 expression).
// ------------------------------------------------------------------------------------------------
__global__ void vit_rope_table_kernel(const int* __restrict__ hpos, const int* __restrict__ wpos,
                                      const float* __restrict__ inv_freq /*[20]*/, float* __restrict__ cosb,
                                      float* __restrict__ sinb, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 40) return;
  const int t = i / 40, j = i % 40;
  const float pos = static_cast<float>(j < 20 ? __ldg(hpos + t) : __ldg(wpos + t));
  const float f = pos * __ldg(inv_freq + (j % 20));
  float s, c;
  sincosf(f, &s, &c);
  cosb[i] = c;
  sinb[i] = s;
}

__global__ void mrope_table_kernel(const int64_t* __restrict__ pos /*[3,S]*/, const float* __restrict__ inv_freq /*[64]*/,
                                   float* __restrict__ cosb, float* __restrict__ sinb, int S, int sec0, int sec1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * 64) return;
  const int t = i >> 6, j = i & 63;
  const int axis = (j < sec0) ? 0 : ((j < sec0 + sec1) ? 1 : 2);
  const float f = static_cast<float>(__ldg(pos + static_cast<int64_t>(axis) * S + t)) * __ldg(inv_freq + j);
  float s, c;
  sincosf(f, &s, &c);
  cosb[i] = c;
  sinb[i] = s;
}

// ------------------------------------------------------------------------------------------------
// ViT: split qkv, apply 2-D RoPE (HF:156-167 apply_rotary_pos_emb_vision, fp32), re-layout for attention:
//   qkv [N, 3*16*80] -> Qp, Kp [N,16,128] (cols 80..127 = 0) and Vt [16*80, ldvt] (transposed).
// Block = 32 tokens x 1 head, 256 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vit_rope_split_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ldqkv,
                                                            const float* __restrict__ cosb, const float* __restrict__ sinb,
                                                            __nv_bfloat16* __restrict__ Qp, __nv_bfloat16* __restrict__ Kp,
                                                            __nv_bfloat16* __restrict__ Vt, int64_t ldvt, int n) {
  constexpr int HD = 80, NH = 16, TOK = 32;
  __shared__ float sq[TOK][HD + 1];
  __shared__ float sk[TOK][HD + 1];
  __shared__ __nv_bfloat16 sv[TOK][HD + 2];
  const int t0 = blockIdx.x * TOK;
  const int h = blockIdx.y;
  const int tid = threadIdx.x;
  // load q, k, v head slices: 32 tokens x 3 x 10 uint4
  for (int i = tid; i < TOK * 30; i += 256) {
    const int t = i / 30, r = i % 30;
    const int which = r / 10, c = r % 10;
    if (t0 + t < n) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(qkv + static_cast<int64_t>(t0 + t) * ldqkv + which * NH * HD + h * HD) + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      if (which == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint32_t*>(&sv[t][c * 8 + j * 2]) = w[j];
      } else {
        float(*dst)[HD + 1] = which == 0 ? sq : sk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dst[t][c * 8 + j * 2] = bf16_lo(w[j]);
          dst[t][c * 8 + j * 2 + 1] = bf16_hi(w[j]);
        }
      }
    }
  }
  __syncthreads();
  // rotate + write Q/K rows: 32 tokens x 2 tensors x 16 uint4 (8 elements each)
  for (int i = tid; i < TOK * 2 * 16; i += 256) {
    const int t = i / 32, r = i % 32;
    const int which = r / 16, c = r % 16;
    if (t0 + t >= n) continue;
    float(*src)[HD + 1] = which == 0 ? sq : sk;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = c * 8 + j;
      float val = 0.f;
      if (d < HD) {
        const int f = d < 40 ? d : d - 40;
        const float cs = __ldg(cosb + static_cast<int64_t>(t0 + t) * 40 + f);
        const float sn = __ldg(sinb + static_cast<int64_t>(t0 + t) * 40 + f);
        const float x = src[t][d];
        const float xr = d < 40 ? -src[t][d + 40] : src[t][d - 40];  // rotate_half
        val = x * cs + xr * sn;
      }
      o[j] = val;
    }
    uint4 pk;
    pk.x = pack_bf16x2(o[0], o[1]);
    pk.y = pack_bf16x2(o[2], o[3]);
    pk.z = pack_bf16x2(o[4], o[5]);
    pk.w = pack_bf16x2(o[6], o[7]);
    __nv_bfloat16* base = (which == 0 ? Qp : Kp) + (static_cast<int64_t>(t0 + t) * NH + h) * 128;
    reinterpret_cast<uint4*>(base)[c] = pk;
  }
  // transposed V: row (h*80 + d), 32 consecutive token columns per warp store
  const int lane = tid & 31, warp = tid >> 5;
  for (int d = warp; d < HD; d += 8) {
    if (t0 + lane < n) Vt[static_cast<int64_t>(h * HD + d) * ldvt + t0 + lane] = sv[lane][d];
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder: M-RoPE on q (in place) and k, KV-cache append (reference mcore .../modeling_qwen2_5_vl.py:21-41;
// HF:627-669, 740-756).  qkv [S, (nq + 2*nkv) * 128]; k -> Kc[cache_row[s], nkv, 128]; v -> Vt[nkv*128, ldvt]
// transposed at column cache_row[s].  Block = 32 tokens, 256 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mrope_kv_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ldqkv,
                                                      const float* __restrict__ cosb, const float* __restrict__ sinb,
                                                      const int* __restrict__ cache_row, __nv_bfloat16* __restrict__ Kc,
                                                      __nv_bfloat16* __restrict__ Vt, int64_t ldvt, int S, int nq, int nkv) {
  // grid = (token blocks of 32, nq + nkv rotary heads + nkv value heads): one (token block, head) per CTA so that the
  // decode step (S = live sequences) spreads over ~20 CTAs instead of serialising 18 heads in one warp.
  constexpr int TOK = 32;
  __shared__ __nv_bfloat16 sv[TOK][128 + 2];
  const int t0 = blockIdx.x * TOK;
  const int hh = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_trigger();
  if (hh < nq + nkv) {
    for (int tt = warp; tt < TOK; tt += 8) {
      const int t = t0 + tt;
      if (t >= S) break;
      const float2 c = __ldg(reinterpret_cast<const float2*>(cosb + static_cast<int64_t>(t) * 64) + lane);
      const float2 sn = __ldg(reinterpret_cast<const float2*>(sinb + static_cast<int64_t>(t) * 64) + lane);
      __nv_bfloat16* row = qkv + static_cast<int64_t>(t) * ldqkv + hh * 128;
      uint32_t* lo = reinterpret_cast<uint32_t*>(row) + lane;
      uint32_t* hi = reinterpret_cast<uint32_t*>(row + 64) + lane;
      const uint32_t a = *lo, b = *hi;
      const float x1a = bf16_lo(a), x1b = bf16_hi(a), x2a = bf16_lo(b), x2b = bf16_hi(b);
      const uint32_t olo = pack_bf16x2(x1a * c.x - x2a * sn.x, x1b * c.y - x2b * sn.y);
      const uint32_t ohi = pack_bf16x2(x2a * c.x + x1a * sn.x, x2b * c.y + x1b * sn.y);
      if (hh < nq) {
        *lo = olo;
        *hi = ohi;
      } else {
        const int crow = __ldg(cache_row + t);
        uint32_t* kd = reinterpret_cast<uint32_t*>(Kc + (static_cast<int64_t>(crow) * nkv + (hh - nq)) * 128);
        kd[lane] = olo;
        kd[32 + lane] = ohi;
      }
    }
  } else {
    const int vh = hh - nq - nkv;
    for (int tt = warp; tt < TOK; tt += 8) {
      const int t = t0 + tt;
      if (t >= S) break;
      const uint32_t* vsrc = reinterpret_cast<const uint32_t*>(qkv + static_cast<int64_t>(t) * ldqkv + (nq + nkv + vh) * 128);
      *reinterpret_cast<uint32_t*>(&sv[tt][lane * 2]) = vsrc[lane];
      *reinterpret_cast<uint32_t*>(&sv[tt][64 + lane * 2]) = vsrc[32 + lane];
    }
    __syncthreads();
    const bool ok = t0 + lane < S;
    const int crow = ok ? __ldg(cache_row + t0 + lane) : 0;
    for (int d = warp; d < 128; d += 8) {
      if (ok) Vt[static_cast<int64_t>(vh * 128 + d) * ldvt + crow] = sv[lane][d];
    }
  }
}

}  // namespace sr

// =================================================================================================
// C ABI
// =================================================================================================
using namespace sr;

extern "C" int sr_rmsnorm_bf16(const void* x, int64_t ldx, const float* weight, void* y, int64_t ldy, int rows, int dim,
                               float eps, void* stream) {
  SR_REQUIRE(x && weight && y, "sr_rmsnorm_bf16: null pointer");
  SR_REQUIRE(rows >= 0 && ldx % 8 == 0 && ldy % 8 == 0, "sr_rmsnorm_bf16: bad rows/strides");
  if (rows == 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (rows + 7) / 8;
  const __nv_bfloat16* xx = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
  if (dim == 1280) rmsnorm_kernel<5><<<grid, 256, 0, s>>>(xx, ldx, weight, yy, ldy, rows, eps);
  else if (dim == 2048) rmsnorm_kernel<8><<<grid, 256, 0, s>>>(xx, ldx, weight, yy, ldy, rows, eps);
  else if (dim == 256) rmsnorm_kernel<1><<<grid, 256, 0, s>>>(xx, ldx, weight, yy, ldy, rows, eps);
  else {
    set_last_error("sr_rmsnorm_bf16: dim %d unsupported (1280 / 2048 / 256)", dim);
    return SR_ERR_UNSUPPORTED;
  }
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_gather_rows(const void* src, int64_t lds, int src_dtype, const int32_t* idx, void* dst, int64_t ldd,
                              int rows, int cols, void* stream) {
  SR_REQUIRE(src && dst, "sr_gather_rows: null pointer");
  SR_REQUIRE(cols % 8 == 0 && ldd % 8 == 0 && lds % (src_dtype == SR_DTYPE_F32 ? 4 : 8) == 0, "sr_gather_rows: cols/strides must be multiples of 8");
  if (rows <= 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (rows + 7) / 8;
  if (src_dtype == SR_DTYPE_F32)
    gather_rows_f32_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(src), lds, idx,
                                                reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols / 8);
  else
    gather_rows_bf16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), lds, idx,
                                                 reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols / 8);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_embed_scatter(const void* table, int64_t ldt, const int64_t* ids, const void* vis, int64_t ldv,
                                const int32_t* vis_idx, void* out, int64_t ldo, int rows, int cols, void* stream) {
  SR_REQUIRE(table && ids && out, "sr_embed_scatter: null pointer");
  SR_REQUIRE(cols % 8 == 0 && ldt % 8 == 0 && ldo % 8 == 0 && (vis == nullptr || ldv % 8 == 0), "sr_embed_scatter: strides");
  if (rows <= 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  embed_scatter_kernel<<<(rows + 7) / 8, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(table), ldt, ids,
                                                     reinterpret_cast<const __nv_bfloat16*>(vis), ldv,
                                                     vis ? vis_idx : nullptr, reinterpret_cast<__nv_bfloat16*>(out), ldo,
                                                     rows, cols / 8);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_vit_rope_table(const int32_t* hpos, const int32_t* wpos, const float* inv_freq, float* cos_out,
                                 float* sin_out, int n, void* stream) {
  SR_REQUIRE(hpos && wpos && inv_freq && cos_out && sin_out, "sr_vit_rope_table: null pointer");
  if (n <= 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  vit_rope_table_kernel<<<(n * 40 + 255) / 256, 256, 0, s>>>(hpos, wpos, inv_freq, cos_out, sin_out, n);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_mrope_table(const int64_t* pos3, const float* inv_freq, float* cos_out, float* sin_out, int S,
                              int sec_t, int sec_h, int sec_w, void* stream) {
  SR_REQUIRE(pos3 && inv_freq && cos_out && sin_out, "sr_mrope_table: null pointer");
  SR_REQUIRE(sec_t + sec_h + sec_w == 64, "sr_mrope_table: sections must sum to 64");
  if (S <= 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  mrope_table_kernel<<<(S * 64 + 255) / 256, 256, 0, s>>>(pos3, inv_freq, cos_out, sin_out, S, sec_t, sec_h);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_vit_rope_split(const void* qkv, int64_t ldqkv, const float* cosb, const float* sinb, void* Qp, void* Kp,
                                 void* Vt, int64_t ldvt, int n, void* stream) {
  SR_REQUIRE(qkv && cosb && sinb && Qp && Kp && Vt, "sr_vit_rope_split: null pointer");
  SR_REQUIRE(ldqkv % 8 == 0 && ldvt >= n, "sr_vit_rope_split: bad strides");
  if (n <= 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((n + 31) / 32, 16);
  vit_rope_split_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), ldqkv, cosb, sinb,
                                             reinterpret_cast<__nv_bfloat16*>(Qp), reinterpret_cast<__nv_bfloat16*>(Kp),
                                             reinterpret_cast<__nv_bfloat16*>(Vt), ldvt, n);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_mrope_kv(void* qkv, int64_t ldqkv, const float* cosb, const float* sinb, const int32_t* cache_row,
                           void* Kc, void* Vt, int64_t ldvt, int S, int n_q_heads, int n_kv_heads, void* stream) {
  SR_REQUIRE(qkv && cosb && sinb && cache_row && Kc && Vt, "sr_mrope_kv: null pointer");
  SR_REQUIRE(ldqkv % 8 == 0 && n_q_heads > 0 && n_kv_heads > 0 && n_kv_heads <= 8, "sr_mrope_kv: bad shape");
  if (S <= 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((S + 31) / 32, n_q_heads + 2 * n_kv_heads);
  SR_CHECK_CUDA(launch_pdl(mrope_kv_kernel, grid, dim3(256), 0, s, reinterpret_cast<__nv_bfloat16*>(qkv), ldqkv, cosb, sinb,
                           cache_row, reinterpret_cast<__nv_bfloat16*>(Kc), reinterpret_cast<__nv_bfloat16*>(Vt), ldvt, S,
                           n_q_heads, n_kv_heads));
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}
