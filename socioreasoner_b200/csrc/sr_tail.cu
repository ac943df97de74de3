// sr_tail.cu -- decode-step attention, LM-head tail (argmax, log-prob / entropy), mask epilogue and IoU.
// All HBM/latency-bound; reported against the HBM roofline.
#include "sr_host.h"

namespace sr {

// ------------------------------------------------------------------------------------------------
// Block reductions (blockDim multiple of 32, <= 1024)
// ------------------------------------------------------------------------------------------------
SR_DEVINL float block_max(float v, float* sbuf) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sbuf[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? sbuf[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}
SR_DEVINL float block_sum(float v, float* sbuf) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sbuf[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? sbuf[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// ------------------------------------------------------------------------------------------------
// Decode-step attention: one new query token per sample against its KV cache (the vLLM decode kernel's job
// in the reference, roll/distributed/strategy/vllm_strategy.py:114-141; HF eager equivalent HF:740-756 with
// use_cache).  q [B, nq, 128] (row stride ldq), Kc [rows, nkv, 128], Vt [nkv*128, ldvt]; sample b attends cache
// rows [kv_start[b], kv_start[b] + kv_len[b]).  One CTA per (sample, q head).
// ------------------------------------------------------------------------------------------------
// Split-KV ("flash-decoding") formulation: CTA = (sample, kv head, key split).  The G = n_q/n_kv query heads that share a
// kv head are processed together, so K and V^T are read ONCE per (sample, kv head) instead of once per query head, and the
// key axis is split across CTAs so that ~all SMs pull cache bytes concurrently.  Per 128-key pass: thread = (key, head
// half) for the scores (16 independent 16-byte loads of the key row in flight), block softmax statistics per head,
// thread = (output dim, head half) for P.V with 16-byte V^T loads; running (max, sum, out) are rescaled between passes.
// Each split writes its un-normalised partial (m, l, o) to the workspace; the last CTA to finish a (sample, kv head)
// (atomic ticket) merges the splits and writes the bf16 output, then re-arms the ticket.
constexpr int DA_THREADS = 256;
constexpr int DA_KEYS = 128;   // keys per pass
constexpr int DA_GMAX = 8;     // max query heads per kv head
__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq,
                                                                const __nv_bfloat16* __restrict__ Kc,
                                                                const __nv_bfloat16* __restrict__ Vt, int64_t ldvt,
                                                                const int* __restrict__ kv_start, const int* __restrict__ kv_len,
                                                                __nv_bfloat16* __restrict__ out, int64_t ldo, int nq, int nkv,
                                                                float scale_log2, float* __restrict__ ws, int* __restrict__ tickets) {
  __shared__ __align__(16) float sq[DA_GMAX][128];
  __shared__ __align__(16) float sp[DA_GMAX][DA_KEYS];   // scores, then probabilities of the current pass
  __shared__ float s_red[8][DA_GMAX];
  __shared__ float s_m[DA_GMAX], s_l[DA_GMAX], s_alpha[DA_GMAX];
  __shared__ int s_last;
  const int b = blockIdx.x, kvh = blockIdx.y, split = blockIdx.z, ns = gridDim.z;
  const int G = nq / nkv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_trigger();
  const int start = __ldg(kv_start + b), len = __ldg(kv_len + b);
  int per = (len + ns - 1) / ns;
  per = (per + 7) & ~7;                      // V^T vectors are 8 keys wide
  const int k_begin = split * per, k_end = min(len, k_begin + per);
  for (int i = tid; i < G * 128; i += DA_THREADS)
    sq[i >> 7][i & 127] = __bfloat162float(q[static_cast<int64_t>(b) * ldq + (kvh * G + (i >> 7)) * 128 + (i & 127)]) * scale_log2;
  if (tid < DA_GMAX) { s_m[tid] = -INFINITY; s_l[tid] = 0.f; }
  const int half = tid >> 7;                 // heads [half*4, half*4+4) (scores) / (P.V)
  const int kk = tid & 127;                  // key within the pass / output dim
  float o_acc[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  for (int p0 = k_begin; p0 < k_end; p0 += DA_KEYS) {
    const int nk = min(DA_KEYS, k_end - p0);
    // ---- scores for (key kk, 4 heads) ----
    float sc4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (kk < nk) {
      const uint4* kr = reinterpret_cast<const uint4*>(Kc + (static_cast<int64_t>(start + p0 + kk) * nkv + kvh) * 128);
      uint4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldg(kr + i);
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const int h = half * 4 + hh;
        if (h < G) {
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 q0 = *reinterpret_cast<const float4*>(&sq[h][i * 8]);
            const float4 q1 = *reinterpret_cast<const float4*>(&sq[h][i * 8 + 4]);
            d += bf16_lo(v[i].x) * q0.x + bf16_hi(v[i].x) * q0.y + bf16_lo(v[i].y) * q0.z + bf16_hi(v[i].y) * q0.w +
                 bf16_lo(v[i].z) * q1.x + bf16_hi(v[i].z) * q1.y + bf16_lo(v[i].w) * q1.z + bf16_hi(v[i].w) * q1.w;
          }
          sc4[hh] = d;
        }
      }
    }
    // ---- per-head max over the pass ----
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      const float m = warp_max(sc4[hh]);
      if (lane == 0) s_red[warp][hh] = m;    // warps 0-3 hold half 0, warps 4-7 half 1
    }
    __syncthreads();
    if (tid < G) {
      const int hf = tid >> 2, hh = tid & 3;
      float m = s_m[tid];
      for (int w = 0; w < 4; ++w) m = fmaxf(m, s_red[hf * 4 + w][hh]);
      const float m_use = (m == -INFINITY) ? 0.f : m;
      s_alpha[tid] = exp2f(s_m[tid] - m_use);   // rescale of the running sum / output (0 on the first pass)
      s_m[tid] = m;
    }
    __syncthreads();
    // ---- probabilities ----
    float ps[4];
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      const int h = half * 4 + hh;
      float pv = 0.f;
      if (h < G && kk < nk) {
        const float m = s_m[h];
        pv = exp2f(sc4[hh] - ((m == -INFINITY) ? 0.f : m));
      }
      if (h < G) sp[h][kk] = pv;             // zero beyond nk: the 8-key vectors below may over-read the tail
      ps[hh] = warp_sum(pv);
      if (lane == 0) s_red[warp][hh] = ps[hh];
    }
    __syncthreads();
    if (tid < G) {
      const int hf = tid >> 2, hh = tid & 3;
      float l = 0.f;
      for (int w = 0; w < 4; ++w) l += s_red[hf * 4 + w][hh];
      s_l[tid] = s_l[tid] * s_alpha[tid] + l;
    }
    // ---- partial O for (dim kk, 4 heads): o = o * alpha + sum_k p[h][k] * Vt[kvh*128 + kk, start + p0 + k] ----
    const uint4* vrow = reinterpret_cast<const uint4*>(Vt + static_cast<int64_t>(kvh * 128 + kk) * ldvt + start + p0);
    const int nvec = (nk + 7) >> 3;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int i = 0; i < nvec; ++i) {
      const uint4 u = __ldg(vrow + i);
      const float vv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const int h = half * 4 + hh;
        if (h < G) {
          const float4 p0v = *reinterpret_cast<const float4*>(&sp[h][i * 8]);
          const float4 p1v = *reinterpret_cast<const float4*>(&sp[h][i * 8 + 4]);
          acc[hh] += vv[0] * p0v.x + vv[1] * p0v.y + vv[2] * p0v.z + vv[3] * p0v.w + vv[4] * p1v.x + vv[5] * p1v.y +
                     vv[6] * p1v.z + vv[7] * p1v.w;
        }
      }
    }
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      const int h = half * 4 + hh;
      if (h < G) o_acc[hh] = o_acc[hh] * s_alpha[h] + acc[hh];
    }
    __syncthreads();   // sp / s_red / s_alpha are reused by the next pass
  }
  // ---- publish this split's partial ----
  const int64_t slot = (static_cast<int64_t>(b) * nkv + kvh) * ns + split;
  float* wo = ws + slot * (DA_GMAX * 128 + 2 * DA_GMAX);
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    const int h = half * 4 + hh;
    if (h < G) wo[h * 128 + kk] = o_acc[hh];
  }
  if (tid < G) {
    wo[DA_GMAX * 128 + tid] = s_m[tid];
    wo[DA_GMAX * 128 + DA_GMAX + tid] = s_l[tid];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int ticket = atomicAdd(tickets + b * nkv + kvh, 1);
    s_last = (ticket == ns - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- merge the splits (last CTA of this (sample, kv head)) ----
  // All loads of the merge are issued in parallel: first the per-split (max, sum) pairs into shared memory, then every
  // thread gathers its ns partial outputs with independent L2 loads (a serial loop here cost ~16 us of L2 round trips).
  constexpr int kStride = DA_GMAX * 128 + 2 * DA_GMAX;
  __shared__ float m_ms[32][DA_GMAX], m_ls[32][DA_GMAX], m_f[32][DA_GMAX];
  __shared__ float m_L[DA_GMAX];
  const float* wb = ws + (static_cast<int64_t>(b) * nkv + kvh) * ns * kStride;
  for (int i = tid; i < ns * G; i += DA_THREADS) {
    const int s2 = i / G, h = i % G;
    m_ms[s2][h] = __ldcg(wb + s2 * kStride + DA_GMAX * 128 + h);
    m_ls[s2][h] = __ldcg(wb + s2 * kStride + DA_GMAX * 128 + DA_GMAX + h);
  }
  __syncthreads();
  if (tid < G) {
    float M = -INFINITY;
    for (int s2 = 0; s2 < ns; ++s2) M = fmaxf(M, m_ms[s2][tid]);
    const float Mu = (M == -INFINITY) ? 0.f : M;
    float L = 0.f;
    for (int s2 = 0; s2 < ns; ++s2) {
      const float f = exp2f(m_ms[s2][tid] - Mu);
      m_f[s2][tid] = f;
      L += m_ls[s2][tid] * f;
    }
    m_L[tid] = L;
  }
  __syncthreads();
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    const int h = half * 4 + hh;
    if (h >= G) continue;
    float O = 0.f;
    for (int s0 = 0; s0 < ns; s0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (s0 + j < ns) ? __ldcg(wb + (s0 + j) * kStride + h * 128 + kk) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) O += (s0 + j < ns) ? v[j] * m_f[s0 + j][h] : 0.f;
    }
    const float L = m_L[h];
    out[static_cast<int64_t>(b) * ldo + (kvh * G + h) * 128 + kk] = __float2bfloat16(L > 0.f ? O / L : 0.f);
  }
  if (tid == 0) tickets[b * nkv + kvh] = 0;   // re-arm for the next launch
}

// ------------------------------------------------------------------------------------------------
// argmax over the vocabulary (greedy decode; first index wins ties, as torch.argmax).
// ------------------------------------------------------------------------------------------------
template <typename T>
SR_DEVINL float to_f32(T v);
template <>
SR_DEVINL float to_f32<float>(float v) { return v; }
template <>
SR_DEVINL float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

constexpr int AM_SPLIT = 64;
SR_DEVINL void argmax_combine(float& best, int& bi, float ov, int oi) {
  if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
}
SR_DEVINL void argmax_block(float& best, int& bi, float* sval, int* sidx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) argmax_combine(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sval[warp] = best; sidx[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < (blockDim.x >> 5) ? sval[lane] : -INFINITY;
    bi = lane < (blockDim.x >> 5) ? sidx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_combine(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
  }
}
// stage 1: grid (rows, AM_SPLIT) -> partial (value, index) per slice; stage 2: grid rows -> final index.
template <typename T>
__global__ void __launch_bounds__(256) argmax_partial_kernel(const T* __restrict__ logits, int64_t ld, int V,
                                                            float* __restrict__ pval, int* __restrict__ pidx) {
  __shared__ float sval[32];
  __shared__ int sidx[32];
  const T* row = logits + static_cast<int64_t>(blockIdx.x) * ld;
  const int per = (V + AM_SPLIT - 1) / AM_SPLIT;
  const int i0 = blockIdx.y * per, i1 = min(V, i0 + per);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) argmax_combine(best, bi, to_f32<T>(row[i]), i);
  argmax_block(best, bi, sval, sidx);
  if (threadIdx.x == 0) {
    pval[blockIdx.x * AM_SPLIT + blockIdx.y] = best;
    pidx[blockIdx.x * AM_SPLIT + blockIdx.y] = bi;
  }
}
__global__ void __launch_bounds__(AM_SPLIT) argmax_final_kernel(const float* __restrict__ pval, const int* __restrict__ pidx,
                                                               int64_t* __restrict__ out) {
  __shared__ float sval[32];
  __shared__ int sidx[32];
  float best = pval[blockIdx.x * AM_SPLIT + threadIdx.x];
  int bi = pidx[blockIdx.x * AM_SPLIT + threadIdx.x];
  argmax_block(best, bi, sval, sidx);
  if (threadIdx.x == 0) out[blockIdx.x] = bi;
}

// ------------------------------------------------------------------------------------------------
// log-prob of the label and entropy per row (roll/utils/functionals.py:193-205):
//   logp = x[label] - logsumexp(x);  H = logsumexp(x) - sum softmax(x) * x.      fp32 math.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) logprob_entropy_kernel(const T* __restrict__ logits, int64_t ld, int V,
                                                              const int64_t* __restrict__ labels,
                                                              float* __restrict__ logprob, float* __restrict__ entropy) {
  __shared__ float red[32];
  const T* row = logits + static_cast<int64_t>(blockIdx.x) * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, to_f32<T>(row[i]));
  mx = block_max(mx, red);
  float se = 0.f, sxe = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float x = to_f32<T>(row[i]);
    const float e = __expf(x - mx);
    se += e;
    sxe += e * x;
  }
  se = block_sum(se, red);
  sxe = block_sum(sxe, red);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(se);
    if (entropy) entropy[blockIdx.x] = lse - sxe / se;
    if (logprob) {
      const int64_t lb = labels ? labels[blockIdx.x] : 0;
      logprob[blockIdx.x] = to_f32<T>(row[lb]) - lse;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Mask epilogue (roll/distributed/strategy/seg_strategy.py:47-65 + the tail of SAM2's predict()):
//   per object: best = argmax(scores[o]) ; bilinear up-sample low-res logits[o,best] (h x w -> image^2,
//   align_corners=False) ; > 0 ; OR over objects ; nearest resize image -> out (cv2.INTER_NEAREST) ; uint8.
// Fused: one thread per OUTPUT pixel evaluates the bilinear sample at its nearest-source pixel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_epilogue_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                                                           int n_obj, int h, int w, int image, int out_size,
                                                           uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= out_size) return;
  // cv2 INTER_NEAREST: sx = min(floor(x * (src/dst)), src-1) with the ratio in double
  const double inv = static_cast<double>(image) / static_cast<double>(out_size);
  int sx = static_cast<int>(floor(x * inv));
  int sy = static_cast<int>(floor(y * inv));
  sx = sx < image - 1 ? sx : image - 1;
  sy = sy < image - 1 ? sy : image - 1;
  // torch upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0; scale = in/out in fp32
  const float scale_h = static_cast<float>(h) / static_cast<float>(image);
  const float scale_w = static_cast<float>(w) / static_cast<float>(image);
  float fy = scale_h * (static_cast<float>(sy) + 0.5f) - 0.5f;
  float fx = scale_w * (static_cast<float>(sx) + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  y0 = y0 < h - 1 ? y0 : h - 1;
  x0 = x0 < w - 1 ? x0 : w - 1;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
  ly = fminf(fmaxf(ly, 0.f), 1.f);
  lx = fminf(fmaxf(lx, 0.f), 1.f);
  const float hy = 1.f - ly, hx = 1.f - lx;
  uint8_t m = 0;
  for (int o = 0; o < n_obj; ++o) {
    const float s0 = __ldg(scores + o * 3), s1 = __ldg(scores + o * 3 + 1), s2 = __ldg(scores + o * 3 + 2);
    int best = 0;
    float bs = s0;
    if (s1 > bs) { bs = s1; best = 1; }
    if (s2 > bs) { bs = s2; best = 2; }
    const float* p = logits + (static_cast<int64_t>(o) * 3 + best) * h * w;
    const float v = hy * (hx * __ldg(p + y0 * w + x0) + lx * __ldg(p + y0 * w + x1)) +
                    ly * (hx * __ldg(p + y1 * w + x0) + lx * __ldg(p + y1 * w + x1));
    m |= (v > 0.f) ? 1 : 0;
  }
  out[static_cast<int64_t>(y) * out_size + x] = m;
}

// ------------------------------------------------------------------------------------------------
// Mask IoU (rlvr_socioseg_vlm_pipeline_infer.py:45-58): sum(p&g)/sum(p|g) on >0 masks, both empty -> empty_value
// (1.0 in the infer pipeline, 0.0 in the reward worker, socioseg_rule_reward_worker.py:181-199).
// One CTA per mask pair, 16-byte loads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mask_iou_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                       int64_t n_pix, float empty_value, float* __restrict__ iou,
                                                       int* __restrict__ counts) {
  __shared__ float red[32];
  const uint8_t* p = pred + static_cast<int64_t>(blockIdx.x) * n_pix;
  const uint8_t* g = gt + static_cast<int64_t>(blockIdx.x) * n_pix;
  int inter = 0, uni = 0;
  const int64_t nvec = n_pix / 16;
  for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p) + i);
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(g) + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // per-byte non-zero -> 0x01 flags
      uint32_t pa = aw[j], pb = bw[j];
      pa = ((pa | (pa >> 4)) & 0x0f0f0f0fu); pa = ((pa | (pa >> 2)) & 0x03030303u); pa = ((pa | (pa >> 1)) & 0x01010101u);
      pb = ((pb | (pb >> 4)) & 0x0f0f0f0fu); pb = ((pb | (pb >> 2)) & 0x03030303u); pb = ((pb | (pb >> 1)) & 0x01010101u);
      inter += __popc(pa & pb);
      uni += __popc(pa | pb);
    }
  }
  for (int64_t i = nvec * 16 + threadIdx.x; i < n_pix; i += blockDim.x) {
    const int a = p[i] > 0, b = g[i] > 0;
    inter += a & b;
    uni += a | b;
  }
  const float fi = block_sum(static_cast<float>(inter), red);
  const float fu = block_sum(static_cast<float>(uni), red);
  if (threadIdx.x == 0) {
    iou[blockIdx.x] = fu > 0.f ? fi / fu : empty_value;
    if (counts) {
      counts[blockIdx.x * 2] = static_cast<int>(fi);
      counts[blockIdx.x * 2 + 1] = static_cast<int>(fu);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Decode-loop bookkeeping on the device (so that one decode step is a replayable CUDA graph): append the token chosen
// by argmax to tokens[:, step] and derive this step's KV-cache slot, KV length and M-RoPE position
// (response positions = last prompt position + 1 + step on all three axes, roll/utils/functionals.py:838-852).
// ------------------------------------------------------------------------------------------------
__global__ void decode_advance_kernel(const int64_t* __restrict__ nxt, int64_t* __restrict__ tokens, int T,
                                      int* __restrict__ step, const int* __restrict__ row0, const int* __restrict__ len0,
                                      const int64_t* __restrict__ pos0, int* __restrict__ cache_row, int* __restrict__ kv_len,
                                      int64_t* __restrict__ pos3, int B) {
  const int s = *step;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (s < T) tokens[static_cast<int64_t>(b) * T + s] = nxt[b];
    cache_row[b] = row0[b] + s;
    kv_len[b] = len0[b] + s + 1;
    const int64_t p = pos0[b] + 1 + s;
    pos3[b] = p;
    pos3[B + b] = p;
    pos3[2 * B + b] = p;
  }
  if (threadIdx.x == 0) *step = s + 1;
}

}  // namespace sr

using namespace sr;

extern "C" int sr_decode_advance(const int64_t* next_tokens, int64_t* tokens, int T, int32_t* step, const int32_t* row0,
                                 const int32_t* len0, const int64_t* pos0, int32_t* cache_row, int32_t* kv_len,
                                 int64_t* pos3, int batch, void* stream) {
  SR_REQUIRE(next_tokens && tokens && step && row0 && len0 && pos0 && cache_row && kv_len && pos3 && batch > 0 && T > 0,
             "sr_decode_advance: bad args");
  decode_advance_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(next_tokens, tokens, T, step, row0, len0, pos0,
                                                                             cache_row, kv_len, pos3, batch);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int64_t sr_decode_attention_workspace_bytes(int batch, int n_kv_heads, int max_len) {
  const int ns = (max_len + DA_KEYS - 1) / DA_KEYS < 1 ? 1 : ((max_len + DA_KEYS - 1) / DA_KEYS > 32 ? 32 : (max_len + DA_KEYS - 1) / DA_KEYS);
  return static_cast<int64_t>(batch) * n_kv_heads * (static_cast<int64_t>(ns) * (DA_GMAX * 128 + 2 * DA_GMAX) * 4 + 16);
}

extern "C" int sr_decode_attention(const void* q, int64_t ldq, const void* Kc, const void* Vt, int64_t ldvt,
                                   const int32_t* kv_start, const int32_t* kv_len, int max_len, void* out, int64_t ldo,
                                   int batch, int n_q_heads, int n_kv_heads, float scale, void* workspace, void* stream) {
  SR_REQUIRE(q && Kc && Vt && kv_start && kv_len && out && workspace, "sr_decode_attention: null pointer");
  SR_REQUIRE(batch > 0 && n_q_heads % n_kv_heads == 0 && n_q_heads / n_kv_heads <= DA_GMAX && max_len > 0,
             "sr_decode_attention: bad shape (query heads per kv head must be <= 8)");
  SR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && ldvt % 8 == 0, "sr_decode_attention: workspace/ldvt alignment");
  int ns = (max_len + DA_KEYS - 1) / DA_KEYS;
  ns = ns < 1 ? 1 : (ns > 32 ? 32 : ns);
  // workspace layout: [batch*n_kv] int32 tickets (zeroed once by the caller, re-armed by the kernel), then partials
  int* tickets = reinterpret_cast<int*>(workspace);
  float* ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((static_cast<size_t>(batch) * n_kv_heads * 4 + 15) / 16) * 16);
  dim3 grid(batch, n_kv_heads, ns);
  SR_CHECK_CUDA(launch_pdl(decode_attn_kernel, grid, dim3(DA_THREADS), 0, reinterpret_cast<cudaStream_t>(stream),
                           reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(Kc),
                           reinterpret_cast<const __nv_bfloat16*>(Vt), ldvt, kv_start, kv_len,
                           reinterpret_cast<__nv_bfloat16*>(out), ldo, n_q_heads, n_kv_heads,
                           scale * 1.4426950408889634f, ws, tickets));
  return SR_OK;
}

extern "C" int sr_argmax_rows(const void* logits, int64_t ld, int dtype, int rows, int V, int64_t* out, void* workspace,
                              void* stream) {
  SR_REQUIRE(logits && out && workspace && rows >= 0 && V > 0, "sr_argmax_rows: bad args (workspace = rows*512 bytes)");
  if (rows == 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* pval = reinterpret_cast<float*>(workspace);
  int* pidx = reinterpret_cast<int*>(pval + static_cast<size_t>(rows) * AM_SPLIT);
  dim3 grid(rows, AM_SPLIT);
  if (dtype == SR_DTYPE_F32)
    argmax_partial_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(logits), ld, V, pval, pidx);
  else
    argmax_partial_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ld, V, pval, pidx);
  argmax_final_kernel<<<rows, AM_SPLIT, 0, s>>>(pval, pidx, out);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_logprob_entropy(const void* logits, int64_t ld, int dtype, int rows, int V, const int64_t* labels,
                                  float* logprob, float* entropy, void* stream) {
  SR_REQUIRE(logits && rows >= 0 && V > 0 && (logprob == nullptr || labels != nullptr), "sr_logprob_entropy: bad args");
  if (rows == 0) return SR_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SR_DTYPE_F32)
    logprob_entropy_kernel<float><<<rows, 1024, 0, s>>>(reinterpret_cast<const float*>(logits), ld, V, labels, logprob, entropy);
  else
    logprob_entropy_kernel<__nv_bfloat16><<<rows, 1024, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ld, V, labels,
                                                                logprob, entropy);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_mask_epilogue(const float* low_res_logits, const float* scores, int n_obj, int h, int w, int image_size,
                                int out_size, uint8_t* out_mask, void* stream) {
  SR_REQUIRE(out_mask && n_obj >= 0 && (n_obj == 0 || (low_res_logits && scores)) && h > 0 && w > 0 && image_size > 0 && out_size > 0,
             "sr_mask_epilogue: bad args");
  dim3 grid((out_size + 255) / 256, out_size);
  mask_epilogue_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(low_res_logits, scores, n_obj, h, w,
                                                                                image_size, out_size, out_mask);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_mask_iou(const uint8_t* pred, const uint8_t* gt, int n_masks, int64_t n_pix, float empty_value, float* iou,
                           int32_t* counts, void* stream) {
  SR_REQUIRE(pred && gt && iou && n_masks >= 0 && n_pix > 0, "sr_mask_iou: bad args");
  SR_REQUIRE(((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0 && n_pix % 16 == 0,
             "sr_mask_iou: masks must be 16B aligned with n_pix %% 16 == 0");
  if (n_masks == 0) return SR_OK;
  mask_iou_kernel<<<n_masks, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, gt, n_pix, empty_value, iou, counts);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

// ------------------------------------------------------------------------------------------------
// Sampling decode (the reference's rollout config: temperature 1.0, top-p 0.8, top-k 100,
// examples/infer/rlvr_megatron.yaml:89-95 -> vLLM SamplingParams, roll/distributed/strategy/vllm_strategy.py:289-309;
// HF order of warpers: temperature -> top-k -> top-p).  One CTA per row:
//   1. exact top-k selection by a 3-round radix select over order-preserving keys (histograms in shared memory),
//   2. candidates (value desc, index asc) bitonic-sorted in shared memory,
//   3. softmax(value / T) over the candidates, nucleus cut: keep j iff exclusive prefix mass < top_p (>= 1 token),
//   4. inverse-CDF pick with the caller-provided uniform u[row, *step]  (deterministic, graph-replayable).
// ------------------------------------------------------------------------------------------------
namespace sr {
constexpr int SMP_THREADS = 1024;
constexpr int SMP_MAXK = 1024;

SR_DEVINL uint32_t order_key(float x) {  // larger float -> larger key
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(SMP_THREADS) sample_topk_topp_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                                      const float* __restrict__ uniforms, int u_ld,
                                                                      const int* __restrict__ step_ptr, float inv_temperature,
                                                                      int top_k, float top_p, int64_t* __restrict__ out) {
  __shared__ int hist[2048];
  __shared__ float cval[SMP_MAXK];
  __shared__ int cidx[SMP_MAXK];
  __shared__ float cprob[SMP_MAXK];
  __shared__ int s_bin, s_need, s_cnt;
  __shared__ float s_red[32];
  const float* row = logits + static_cast<int64_t>(blockIdx.x) * ld;
  const int tid = threadIdx.x;
  const int k = min(top_k, min(V, SMP_MAXK));
  // ---- radix select: find the key of the k-th largest element ----
  uint32_t prefix = 0, prefix_mask = 0;
  int need = k;  // how many elements with (key & mask) == prefix we still need from the current bucket
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
  for (int r = 0; r < 3; ++r) {
    const int nb = 1 << bits[r];
    for (int i = tid; i < nb; i += SMP_THREADS) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < V; i += SMP_THREADS) {
      const uint32_t key = order_key(row[i]);
      if ((key & prefix_mask) == prefix) atomicAdd(&hist[(key >> shifts[r]) & (nb - 1)], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int acc = 0, b = nb - 1;
      for (; b >= 0; --b) {
        if (acc + hist[b] >= need) break;
        acc += hist[b];
      }
      s_bin = b;
      s_need = need - acc;
    }
    __syncthreads();
    prefix |= static_cast<uint32_t>(s_bin) << shifts[r];
    prefix_mask |= static_cast<uint32_t>(nb - 1) << shifts[r];
    need = s_need;
    __syncthreads();
  }
  const uint32_t kth_key = prefix;  // exact key of the k-th largest; `need` of the elements equal to it are taken (lowest index first)
  // ---- collect candidates: all keys > kth_key, plus the `need` lowest-index elements equal to it ----
  __shared__ int tie_idx[SMP_MAXK];
  __shared__ int s_ties;
  if (tid == 0) { s_cnt = 0; s_ties = 0; }
  for (int i = tid; i < SMP_MAXK; i += SMP_THREADS) { cval[i] = -INFINITY; cidx[i] = 0x7fffffff; }
  __syncthreads();
  for (int i = tid; i < V; i += SMP_THREADS) {
    const float x = row[i];
    const uint32_t key = order_key(x);
    if (key > kth_key) {
      const int slot = atomicAdd(&s_cnt, 1);
      cval[slot] = x;
      cidx[slot] = i;
    } else if (key == kth_key) {
      const int slot = atomicAdd(&s_ties, 1);
      if (slot < SMP_MAXK) tie_idx[slot] = i;  // > 1024 exact ties at the threshold: any 1024 of them (degenerate logits)
    }
  }
  __syncthreads();
  const int n_gt = s_cnt;
  if (tid == 0) {
    const int nt = min(s_ties, SMP_MAXK);
    for (int t2 = 0; t2 < need; ++t2) {  // `need` smallest indices among the ties (nt is 1-2 in practice)
      int best = -1;
      for (int j = 0; j < nt; ++j)
        if (tie_idx[j] != 0x7fffffff && (best < 0 || tie_idx[j] < tie_idx[best])) best = j;
      if (best < 0) break;
      cval[n_gt + t2] = row[tie_idx[best]];
      cidx[n_gt + t2] = tie_idx[best];
      tie_idx[best] = 0x7fffffff;
    }
  }
  __syncthreads();
  // ---- bitonic sort of SMP_MAXK (value desc, index asc); padding = (-inf, INT_MAX) sinks to the end ----
  for (int size = 2; size <= SMP_MAXK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int i = tid, j = i ^ stride;
      if (j > i) {
        const bool desc = (i & size) == 0;
        const float vi = cval[i], vj = cval[j];
        const int ii = cidx[i], ij = cidx[j];
        const bool i_before_j = (vi > vj) || (vi == vj && ii < ij);
        if (desc ? !i_before_j : i_before_j) {
          cval[i] = vj; cval[j] = vi;
          cidx[i] = ij; cidx[j] = ii;
        }
      }
      __syncthreads();
    }
  }
  // ---- softmax over the k candidates at temperature T ----
  const float vmax = cval[0];
  const float e = (tid < k) ? __expf((cval[tid] - vmax) * inv_temperature) : 0.f;
  const float tot = block_sum(e, s_red);
  cprob[tid] = e / tot;
  __syncthreads();
  // inclusive prefix sum (Hillis-Steele over 1024 entries)
  for (int off = 1; off < SMP_MAXK; off <<= 1) {
    const float add = (tid >= off) ? cprob[tid - off] : 0.f;
    __syncthreads();
    cprob[tid] += add;
    __syncthreads();
  }
  // nucleus: keep j iff exclusive prefix < top_p  (j = 0 always kept)
  const float incl = cprob[tid];
  const float p_j = e / tot;
  const bool keep = (tid < k) && (tid == 0 || (incl - p_j) < top_p);
  // kept set is a prefix [0, n_keep)
  const float kept_mass = block_sum(keep ? p_j : 0.f, s_red);
  const int n_keep = static_cast<int>(block_sum(keep ? 1.f : 0.f, s_red) + 0.5f);
  const float u = uniforms[static_cast<int64_t>(blockIdx.x) * u_ld + (step_ptr ? *step_ptr : 0)] * kept_mass;
  // pick the first kept j with inclusive prefix > u  (u == kept mass numerically -> last kept)
  __shared__ int s_pick;
  if (tid == 0) s_pick = 0x7fffffff;
  __syncthreads();
  if (keep && incl > u) atomicMin(&s_pick, tid);
  __syncthreads();
  if (tid == 0) out[blockIdx.x] = cidx[min(s_pick, n_keep - 1)];
}
}  // namespace sr

extern "C" int sr_sample_topk_topp(const float* logits, int64_t ld, int rows, int V, const float* uniforms, int u_ld,
                                   const int32_t* step, float temperature, int top_k, float top_p, int64_t* out, void* stream) {
  SR_REQUIRE(logits && uniforms && out && rows >= 0 && V > 0, "sr_sample_topk_topp: bad args");
  SR_REQUIRE(temperature > 0.f && top_k >= 1 && top_k <= sr::SMP_MAXK && top_p > 0.f && top_p <= 1.f,
             "sr_sample_topk_topp: need temperature > 0, 1 <= top_k <= 1024, 0 < top_p <= 1");
  if (rows == 0) return SR_OK;
  sr::sample_topk_topp_kernel<<<rows, sr::SMP_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, ld, V, uniforms, u_ld, step, 1.0f / temperature, top_k, top_p, out);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}
