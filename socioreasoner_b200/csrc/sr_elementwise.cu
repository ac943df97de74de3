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
//   qkv [N, 3*16*80] -> Qp, Kp [N,16,128] (columns 0..79; 80..127 are zero padding that the caller zero-fills ONCE --
//   this kernel never writes them) and Vt [16*80, ldvt] (transposed).
// Block = 32 tokens x all heads, 256 threads.  rotate_half pairs element d with d+40, i.e. 16-byte group j (j < 5)
// with group j+5: one work item = (token, q|k, head, j) = 2 vector loads -> 2 vector stores, no staging of q/k;
// cos/sin of the 32 tokens are staged in shared memory once (each value is reused 32x), V goes through a padded
// shared-memory tile for the transpose.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vit_rope_split_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ldqkv,
                                                            const float* __restrict__ cosb, const float* __restrict__ sinb,
                                                            __nv_bfloat16* __restrict__ Qp, __nv_bfloat16* __restrict__ Kp,
                                                            __nv_bfloat16* __restrict__ Vt, int64_t ldvt, int n) {
  constexpr int HD = 80, NH = 16, TOK = 32, VH = 4;  // VH heads of V per transpose pass
  __shared__ __align__(16) float sc[TOK][40];
  __shared__ __align__(16) float ss[TOK][40];
  __shared__ __nv_bfloat16 sv[TOK][VH * HD + 2];
  const int t0 = blockIdx.x * TOK;
  const int tid = threadIdx.x;
  const int ntok = min(TOK, n - t0);
  for (int i = tid; i < ntok * 10; i += 256) {  // 40 floats = 10 float4 per token
    const int t = i / 10, c = i % 10;
    reinterpret_cast<float4*>(&sc[t][0])[c] = __ldg(reinterpret_cast<const float4*>(cosb + static_cast<int64_t>(t0 + t) * 40) + c);
    reinterpret_cast<float4*>(&ss[t][0])[c] = __ldg(reinterpret_cast<const float4*>(sinb + static_cast<int64_t>(t0 + t) * 40) + c);
  }
  __syncthreads();
  // q / k rotation: items = ntok * 2 * 16 * 5
  const int items = ntok * 160;
#pragma unroll 2
  for (int i = tid; i < items; i += 256) {
    const int t = i / 160, r = i % 160;
    const int which = r / 80, hj = r % 80;
    const int h = hj / 5, j = hj % 5;
    const __nv_bfloat16* src = qkv + static_cast<int64_t>(t0 + t) * ldqkv + which * NH * HD + h * HD + j * 8;
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(src));
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(src + 40));
    const float4 c0 = *reinterpret_cast<const float4*>(&sc[t][j * 8]), c1 = *reinterpret_cast<const float4*>(&sc[t][j * 8 + 4]);
    const float4 s0 = *reinterpret_cast<const float4*>(&ss[t][j * 8]), s1 = *reinterpret_cast<const float4*>(&ss[t][j * 8 + 4]);
    const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t ol[4], oh[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x1a = bf16_lo(lw[e]), x1b = bf16_hi(lw[e]), x2a = bf16_lo(hw[e]), x2b = bf16_hi(hw[e]);
      // out[d] = x[d] cos - x[d+40] sin ; out[d+40] = x[d+40] cos + x[d] sin   (cos/sin of d and d+40 are identical)
      ol[e] = pack_bf16x2(x1a * cs[2 * e] - x2a * sn[2 * e], x1b * cs[2 * e + 1] - x2b * sn[2 * e + 1]);
      oh[e] = pack_bf16x2(x2a * cs[2 * e] + x1a * sn[2 * e], x2b * cs[2 * e + 1] + x1b * sn[2 * e + 1]);
    }
    __nv_bfloat16* dst = (which == 0 ? Qp : Kp) + (static_cast<int64_t>(t0 + t) * NH + h) * 128 + j * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    *reinterpret_cast<uint4*>(dst + 40) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
  }
  // transposed V, VH heads per pass: row (h*80 + d) of Vt gets 32 consecutive token columns per warp store
  const int lane = tid & 31, warp = tid >> 5;
  for (int h0 = 0; h0 < NH; h0 += VH) {
    __syncthreads();
    for (int i = tid; i < ntok * (VH * HD / 8); i += 256) {
      const int t = i / (VH * HD / 8), c = i % (VH * HD / 8);
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(qkv + static_cast<int64_t>(t0 + t) * ldqkv + 2 * NH * HD + h0 * HD) + c);
      uint32_t* d = reinterpret_cast<uint32_t*>(&sv[t][c * 8]);
      d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
    }
    __syncthreads();
    for (int d = warp; d < VH * HD; d += 8) {
      if (lane < ntok) Vt[static_cast<int64_t>(h0 * HD + d) * ldvt + t0 + lane] = sv[lane][d];
    }
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
  dim3 grid((n + 31) / 32);
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
