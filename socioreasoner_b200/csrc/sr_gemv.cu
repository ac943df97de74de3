// sr_gemv.cu -- skinny GEMM for the decode step: C[M<=32, N] = epilogue(A[M,K] . W[N,K]^T), M = live sequences.
//
// One new token per sequence makes every linear layer of the decoder (HF modeling_qwen2_5_vl.py:704-707, 611-624,
// 1385; the reference runs them inside vLLM's decode loop, roll/distributed/strategy/vllm_strategy.py:114-141) a
// weight-streaming problem: 2 bytes of W per 2*M flops -> HBM-bound for M <= ~150.  The job of this kernel is to keep
// ~6 TB/s of weight bytes in flight, not to feed tcgen05 (a 128-row UMMA tile would be 94% padding):
//   * CTA = 16 output features (32 packed rows for SwiGLU) x the whole K; its 8 warps split K in interleaved
//     64-byte chunks, so the CTA reads 512 contiguous bytes per weight row per iteration;
//   * every thread streams 16-byte vectors of W straight into registers (ld.global.nc, L1 no-allocate), 4 chunks
//     (8..16 independent 16 B loads) in flight, and feeds them to mma.sync.m16n8k16 (bf16 -> fp32) as the A operand
//     with the k index permuted so that one 16 B load IS an A fragment (the activation fragment uses the same
//     permutation, so no shuffles / shared-memory staging of W);
//   * the activations (M x K bf16, L2-resident) are read with the same 16 B pattern as the B operand;
//   * cross-warp K reduction through 4 KB of shared memory, then the fused epilogue (bias / residual / SwiGLU).
// Reported against the HBM roofline (algorithmic bytes = N*K*2).
#include <stdlib.h>

#include "sr_host.h"

namespace sr {



// weight-stream load flavours (tuning knob sr_set_gemv_load): 0 = nc + L1 no-allocate, 1 = the same with a 256-byte L2
// prefetch granule (bigger DRAM bursts on the linear streaming layout), 2 = plain cached load
template <int LDV>
SR_DEVINL uint4 ld_stream_v4(const void* p) {
  uint4 r;
  if (LDV == 1)
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if (LDV == 2)
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

SR_DEVINL void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct GemvParams {
  const __nv_bfloat16* A;
  const __nv_bfloat16* W;
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  int64_t lda, ldw, ldc, ldr;
  int M, N, K;
  int ksplit;           // 1, or 2: two CTAs per row group split K; partials meet in `ws` (ticket in `tickets`), NB=1 / no fused norm
  float* ws;
  int* tickets;
  int packed;           // 1: W is in the streaming layout [N/16][K/32][16 rows][32 k] (1 KB blocks, see sr_pack_skinny_weights)
  const float* norm_w;  // optional fused RMSNorm on A rows: A'[m,k] = A[m,k] * rsqrt(mean_k A^2 + eps) * norm_w[k]
  float norm_eps;
};

// RG = row groups of 16 weight rows per CTA (1, or 2 for SwiGLU: gate group + up group); NB = batch tiles of 8.
SR_DEVINL uint32_t scale_bf16x2(uint32_t x, float w0, float w1) { return pack_bf16x2(bf16_lo(x) * w0, bf16_hi(x) * w1); }

template <int RG, int NB, int EPI, bool OUT_F32, int GV_WARPS, bool NORM, int GV_UNROLL, int LDV>
__global__ void __launch_bounds__(GV_WARPS * 32) gemv_kernel(const GemvParams p) {
  constexpr int GV_THREADS = GV_WARPS * 32;
  __shared__ float red[GV_WARPS][RG][NB][16 * 8];
  __shared__ float ssq_red[GV_WARPS][NB][8];
  __shared__ float inv_rms[NB][8];
  float ssq[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) ssq[n] = 0.f;
  pdl_trigger();  // let the next decode kernel become resident; it blocks in its own pdl_wait() until we are done
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * (16 * RG);
  float acc[RG][NB][4];
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int n = 0; n < NB; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][n][i] = 0.f;

  // row pointers (clamped: rows >= N are read but never stored)
  const __nv_bfloat16* wrow[RG][2];
#pragma unroll
  for (int r = 0; r < RG; ++r) {
    int ra = row0 + r * 16 + g, rb = ra + 8;
    ra = ra < p.N ? ra : p.N - 1;
    rb = rb < p.N ? rb : p.N - 1;
    if (p.packed) {
      // block (row group, chunk) = 16 rows x 32 k contiguous (1 KB): a warp-wide 16 B load covers 8 rows x 64 B = 512
      // contiguous bytes (4 full 128 B lines instead of 8 half-used ones) and the CTA streams one linear region
      const int64_t grp = static_cast<int64_t>(row0 / 16 + r) * (p.K >> 5) * 512;
      wrow[r][0] = p.W + grp + g * 32 + t * 8;
      wrow[r][1] = p.W + grp + (g + 8) * 32 + t * 8;
    } else {
      wrow[r][0] = p.W + static_cast<int64_t>(ra) * p.ldw + t * 8;
      wrow[r][1] = p.W + static_cast<int64_t>(rb) * p.ldw + t * 8;
    }
  }
  const __nv_bfloat16* xrow[NB];
  bool xok[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    const int m = n * 8 + g;
    xok[n] = m < p.M;
    xrow[n] = p.A + static_cast<int64_t>(xok[n] ? m : 0) * p.lda + t * 8;
  }

  const int nchunks_all = p.K >> 5;  // 32 k per chunk
  const int cper = (nchunks_all + p.ksplit - 1) / p.ksplit;
  const int c_begin = blockIdx.y * cper;
  const int nchunks = min(nchunks_all, c_begin + cper);  // this CTA handles chunks [c_begin, nchunks)
  const int wstep = p.packed ? 512 : 32;  // elements between consecutive chunks of one row group
  // Rolling software pipeline over GV_UNROLL register slots: slot u is consumed (MMA) and immediately re-armed with the
  // chunk GV_WARPS*GV_UNROLL further on, so GV_UNROLL-1 chunk loads per thread stay in flight across iterations
  // (without this the kernel paid one full HBM latency per batch of loads).
  uint4 w[GV_UNROLL][RG][2];
  uint4 x[GV_UNROLL][NB];
  float4 nw0[GV_UNROLL], nw1[GV_UNROLL];
  auto load_w = [&](int u, int c) {
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      w[u][r][0] = ld_stream_v4<LDV>(wrow[r][0] + static_cast<int64_t>(c) * wstep);
      w[u][r][1] = ld_stream_v4<LDV>(wrow[r][1] + static_cast<int64_t>(c) * wstep);
    }
  };
  auto load_x = [&](int u, int c) {
#pragma unroll
    for (int n = 0; n < NB; ++n) x[u][n] = xok[n] ? *reinterpret_cast<const uint4*>(xrow[n] + c * 32) : make_uint4(0, 0, 0, 0);
    if (NORM) {
      nw0[u] = __ldg(reinterpret_cast<const float4*>(p.norm_w + c * 32 + t * 8));
      nw1[u] = __ldg(reinterpret_cast<const float4*>(p.norm_w + c * 32 + t * 8 + 4));
    }
  };
  auto compute_slot = [&](int u) {
    if (NORM) {
      // fused RMSNorm (HF:57-71): the activation fragment is multiplied by the norm weight here; sum(x^2) is
      // accumulated on the side and 1/rms is applied to the fp32 result (every CTA reads all of A anyway)
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const uint4 xv = x[u][n];
        const float a0 = bf16_lo(xv.x), a1 = bf16_hi(xv.x), a2 = bf16_lo(xv.y), a3 = bf16_hi(xv.y);
        const float a4 = bf16_lo(xv.z), a5 = bf16_hi(xv.z), a6 = bf16_lo(xv.w), a7 = bf16_hi(xv.w);
        ssq[n] += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7;
        x[u][n].x = pack_bf16x2(a0 * nw0[u].x, a1 * nw0[u].y);
        x[u][n].y = pack_bf16x2(a2 * nw0[u].z, a3 * nw0[u].w);
        x[u][n].z = pack_bf16x2(a4 * nw1[u].x, a5 * nw1[u].y);
        x[u][n].w = pack_bf16x2(a6 * nw1[u].z, a7 * nw1[u].w);
      }
    }
#pragma unroll
    for (int r = 0; r < RG; ++r)
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        // slots (2t,2t+1 | 2t+8,2t+9) <-> actual k = c*32 + t*8 + {0,1 | 2,3}, then {4,5 | 6,7}
        mma_bf16_16816(acc[r][n], w[u][r][0].x, w[u][r][1].x, w[u][r][0].y, w[u][r][1].y, x[u][n].x, x[u][n].y);
        mma_bf16_16816(acc[r][n], w[u][r][0].z, w[u][r][1].z, w[u][r][0].w, w[u][r][1].w, x[u][n].z, x[u][n].w);
      }
  };
#pragma unroll
  // The first wave of WEIGHT loads does not depend on the previous kernel: issue it before the programmatic-dependency
  // wait so the HBM stream of this layer starts while the previous decode kernel is still draining.
  for (int u = 0; u < GV_UNROLL; ++u) {
    const int c = c_begin + warp + u * GV_WARPS;
    if (c < nchunks) load_w(u, c);
  }
  pdl_wait();  // activations / residual / output buffers belong to the previous kernel until here
#pragma unroll
  for (int u = 0; u < GV_UNROLL; ++u) {
    const int c = c_begin + warp + u * GV_WARPS;
    if (c < nchunks) load_x(u, c);
  }
  for (int base = c_begin + warp; base < nchunks; base += GV_WARPS * GV_UNROLL) {
#pragma unroll
    for (int u = 0; u < GV_UNROLL; ++u) {
      const int c = base + u * GV_WARPS;
      if (c < nchunks) {
        compute_slot(u);
        const int cn = c + GV_WARPS * GV_UNROLL;
        if (cn < nchunks) { load_w(u, cn); load_x(u, cn); }
      }
    }
  }
  // cross-warp reduction: C fragment (feature g | g+8, batch 2t | 2t+1) -> red[warp][r][n][feature*8 + batch]
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      red[warp][r][n][g * 8 + 2 * t] = acc[r][n][0];
      red[warp][r][n][g * 8 + 2 * t + 1] = acc[r][n][1];
      red[warp][r][n][(g + 8) * 8 + 2 * t] = acc[r][n][2];
      red[warp][r][n][(g + 8) * 8 + 2 * t + 1] = acc[r][n][3];
    }
  if (NORM) {
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      float v = ssq[n];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0) ssq_red[warp][n][g] = v;
    }
  }
  __syncthreads();
  if (NORM) {
    if (threadIdx.x < NB * 8) {
      const int n = threadIdx.x >> 3, r = threadIdx.x & 7;
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GV_WARPS; ++w8) v += ssq_red[w8][n][r];
      inv_rms[n][r] = rsqrtf(v / static_cast<float>(p.K) + p.norm_eps);
    }
    __syncthreads();
  }
  // split-K: the two CTAs of a row group exchange fp32 partials through global memory; the second to arrive (ticket)
  // adds its partner's partial and runs the epilogue, the first one leaves
  const bool use_split = (NB == 1 && RG == 1 && !NORM && p.ksplit == 2);
  float vsplit = 0.f;
  if (use_split) {
    __shared__ int s_ticket;
    float* myws = p.ws + (static_cast<int64_t>(blockIdx.x) * 2 + blockIdx.y) * 128;
    if (threadIdx.x < 128) {
      const int b = threadIdx.x >> 4, f = threadIdx.x & 15;
      float sacc = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GV_WARPS; ++w8) sacc += red[w8][0][0][f * 8 + b];
      vsplit = sacc;
      myws[threadIdx.x] = sacc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.tickets + blockIdx.x, 1);
    __syncthreads();
    if (s_ticket == 0) return;
    __threadfence();
    if (threadIdx.x < 128)
      vsplit += __ldcg(p.ws + (static_cast<int64_t>(blockIdx.x) * 2 + (1 - blockIdx.y)) * 128 + threadIdx.x);
    if (threadIdx.x == 0) p.tickets[blockIdx.x] = 0;  // re-arm
  }
  // epilogue: thread -> (batch tile n, feature f, batch b); 128 outputs per batch tile
  for (int idx = threadIdx.x; idx < NB * 128; idx += GV_THREADS) {
    const int n = idx >> 7, e = idx & 127;
    const int b = e >> 4, f = e & 15;  // consecutive threads -> consecutive features (contiguous stores)
    const int m = n * 8 + b;
    float v[RG];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GV_WARPS; ++w8) s += red[w8][r][n][f * 8 + b];
      v[r] = use_split ? vsplit : (NORM ? s * inv_rms[n][b] : s);
    }
    if (m >= p.M) continue;
    if (EPI == SR_EPI_SWIGLU) {
      const int prow = row0 + f;  // packed gate row; up row = prow + 16
      if (prow + 16 >= p.N) continue;
      float gate = v[0], up = v[RG - 1];
      if (p.bias) { gate += __ldg(p.bias + prow); up += __ldg(p.bias + prow + 16); }
      const float o = (gate / (1.0f + __expf(-gate))) * up;
      const int oc = (row0 >> 1) + f;
      reinterpret_cast<__nv_bfloat16*>(p.C)[static_cast<int64_t>(m) * p.ldc + oc] = __float2bfloat16(o);
    } else {
      const int col = row0 + f;
      if (col >= p.N) continue;
      float o = v[0];
      if (p.bias) o += __ldg(p.bias + col);
      if (EPI == SR_EPI_GELU) o = 0.5f * o * (1.0f + erff(o * 0.70710678118654752440f));
      if (EPI == SR_EPI_RESIDUAL) o += __bfloat162float(p.residual[static_cast<int64_t>(m) * p.ldr + col]);
      if (OUT_F32) reinterpret_cast<float*>(p.C)[static_cast<int64_t>(m) * p.ldc + col] = o;
      else reinterpret_cast<__nv_bfloat16*>(p.C)[static_cast<int64_t>(m) * p.ldc + col] = __float2bfloat16(o);
    }
  }
}

static int g_gemv_load = 0;   // weight-load flavour for M <= 8 (see ld_stream_v4)
static int g_gemv_warps = 0;  // 0 = default policy; 4 / 8 / 16 force (tuning knob, see sr_set_gemv_warps)

template <int NB, bool NORM, int W, int LDV>
static int launch_gemv_w(const GemvParams& p, int epi, int out_f32, cudaStream_t s) {
  constexpr int GV_THREADS = W * 32;
  constexpr int U1 = (NB == 1) ? (NORM ? 4 : 6) : 4;  // register slots (chunks in flight + 1) per thread, 1 row group
  constexpr int U2 = (NB == 1) ? 3 : 2;               // 2 row groups (SwiGLU) load twice the weight bytes per slot
  if (epi == SR_EPI_SWIGLU) {
    const int grid = (p.N + 31) / 32;
    SR_CHECK_CUDA(launch_pdl(gemv_kernel<2, NB, SR_EPI_SWIGLU, false, W, NORM, U2, LDV>, dim3(grid), dim3(GV_THREADS), 0, s, p));
  } else {
    const dim3 grid((p.N + 15) / 16, p.ksplit);
    if (out_f32) SR_CHECK_CUDA(launch_pdl(gemv_kernel<1, NB, SR_EPI_NONE, true, W, NORM, U1, LDV>, dim3(grid), dim3(GV_THREADS), 0, s, p));
    else if (epi == SR_EPI_NONE) SR_CHECK_CUDA(launch_pdl(gemv_kernel<1, NB, SR_EPI_NONE, false, W, NORM, U1, LDV>, dim3(grid), dim3(GV_THREADS), 0, s, p));
    else if (epi == SR_EPI_RESIDUAL) SR_CHECK_CUDA(launch_pdl(gemv_kernel<1, NB, SR_EPI_RESIDUAL, false, W, NORM, U1, LDV>, dim3(grid), dim3(GV_THREADS), 0, s, p));
    else if (epi == SR_EPI_GELU) SR_CHECK_CUDA(launch_pdl(gemv_kernel<1, NB, SR_EPI_GELU, false, W, NORM, U1, LDV>, dim3(grid), dim3(GV_THREADS), 0, s, p));
    else {
      set_last_error("sr_gemm_skinny_bf16: unknown epilogue %d", epi);
      return SR_ERR_INVALID;
    }
  }
  return SR_OK;
}

template <int NB, bool NORM>
static int launch_gemv(const GemvParams& p, int epi, int out_f32, cudaStream_t s) {
  // Warps per CTA split K.  Fewer warps per CTA = more CTAs co-resident per SM, so one CTA's reduction/epilogue (no loads
  // in flight) overlaps another CTA's streaming; more warps = shorter per-CTA chains.  Shared-memory budget caps NB >= 2 at 8.
  if constexpr (NB == 1) {
    // measured on B200 (tools/gemv_bench.py, cold weights): 16 warps for the 16-row groups (N = 2048 gives only 128 CTAs:
    // long K needs the wide K split), 8 warps for SwiGLU pairs (688+ CTAs: two co-resident CTAs overlap epilogue and streaming)
    const int w = g_gemv_warps ? g_gemv_warps : (epi == SR_EPI_SWIGLU ? 8 : 16);
    if (g_gemv_load == 1) {
      if (w == 16) return launch_gemv_w<NB, NORM, 16, 1>(p, epi, out_f32, s);
      return launch_gemv_w<NB, NORM, 8, 1>(p, epi, out_f32, s);
    }
    if (g_gemv_load == 2) {
      if (w == 16) return launch_gemv_w<NB, NORM, 16, 2>(p, epi, out_f32, s);
      return launch_gemv_w<NB, NORM, 8, 2>(p, epi, out_f32, s);
    }
    if (w == 4) return launch_gemv_w<NB, NORM, 4, 0>(p, epi, out_f32, s);
    if (w == 16) return launch_gemv_w<NB, NORM, 16, 0>(p, epi, out_f32, s);
    return launch_gemv_w<NB, NORM, 8, 0>(p, epi, out_f32, s);
  } else {
    return launch_gemv_w<NB, NORM, 8, 0>(p, epi, out_f32, s);
  }
}

}  // namespace sr

static int gemm_skinny_impl(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, const float* bias,
                            const void* residual, int64_t ldr, int M, int N, int K, int epilogue, int out_dtype,
                            const float* norm_weight, float norm_eps, void* stream, int packed = 0, int ksplit = 1,
                            void* workspace = nullptr) {
  using namespace sr;
  SR_REQUIRE(A && W && C, "sr_gemm_skinny_bf16: null operand");
  SR_REQUIRE(M > 0 && M <= 32 && N > 0 && K > 0 && K % 32 == 0, "sr_gemm_skinny_bf16: need 0 < M <= 32, K %% 32 == 0 (M=%d K=%d)", M, K);
  SR_REQUIRE(lda % 8 == 0 && (packed || ldw % 8 == 0), "sr_gemm_skinny_bf16: lda/ldw must be multiples of 8");
  SR_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "sr_gemm_skinny_bf16: A/W not 16B aligned");
  SR_REQUIRE(epilogue != SR_EPI_RESIDUAL || residual, "sr_gemm_skinny_bf16: residual epilogue needs residual");
  SR_REQUIRE(epilogue != SR_EPI_SWIGLU || (N % 32 == 0 && out_dtype == SR_DTYPE_BF16), "sr_gemm_skinny_bf16: SwiGLU needs N %% 32 == 0, bf16 out");
  SR_REQUIRE(out_dtype == SR_DTYPE_BF16 || epilogue == SR_EPI_NONE, "sr_gemm_skinny_bf16: fp32 output supports SR_EPI_NONE only");
  GemvParams p;
  p.A = reinterpret_cast<const __nv_bfloat16*>(A);
  p.W = reinterpret_cast<const __nv_bfloat16*>(W);
  p.C = C;
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.lda = lda; p.ldw = ldw; p.ldc = ldc; p.ldr = ldr;
  p.M = M; p.N = N; p.K = K;
  p.norm_w = norm_weight;
  p.norm_eps = norm_eps;
  p.packed = packed;
  p.ksplit = 1;
  p.ws = nullptr;
  p.tickets = nullptr;
  if (ksplit == 2 && M <= 8 && norm_weight == nullptr && epilogue != SR_EPI_SWIGLU) {
    SR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sr_gemm_skinny split-K: workspace missing / unaligned");
    const int groups = (N + 15) / 16;
    p.ksplit = 2;
    p.tickets = reinterpret_cast<int*>(workspace);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((static_cast<size_t>(groups) * 4 + 15) / 16) * 16);
  }
  SR_REQUIRE(!packed || N % 16 == 0, "sr_gemm_skinny_packed_bf16: N must be a multiple of 16");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int f32 = out_dtype == SR_DTYPE_F32;
  if (norm_weight != nullptr) {
    SR_REQUIRE((reinterpret_cast<uintptr_t>(norm_weight) & 15) == 0, "sr_gemm_skinny_norm_bf16: norm_weight not 16B aligned");
    if (M <= 8) return launch_gemv<1, true>(p, epilogue, f32, s);
    if (M <= 16) return launch_gemv<2, true>(p, epilogue, f32, s);
    return launch_gemv<4, true>(p, epilogue, f32, s);
  }
  if (M <= 8) return launch_gemv<1, false>(p, epilogue, f32, s);
  if (M <= 16) return launch_gemv<2, false>(p, epilogue, f32, s);
  return launch_gemv<4, false>(p, epilogue, f32, s);
}

extern "C" int sr_gemm_skinny_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                                   const float* bias, const void* residual, int64_t ldr, int M, int N, int K, int epilogue,
                                   int out_dtype, void* stream) {
  return gemm_skinny_impl(A, lda, W, ldw, C, ldc, bias, residual, ldr, M, N, K, epilogue, out_dtype, nullptr, 0.f, stream);
}

extern "C" int sr_gemm_skinny_norm_bf16(const void* A, int64_t lda, const float* norm_weight, float norm_eps, const void* W,
                                        int64_t ldw, void* C, int64_t ldc, const float* bias, const void* residual,
                                        int64_t ldr, int M, int N, int K, int epilogue, int out_dtype, void* stream) {
  using namespace sr;
  SR_REQUIRE(norm_weight != nullptr, "sr_gemm_skinny_norm_bf16: norm_weight is null");
  return gemm_skinny_impl(A, lda, W, ldw, C, ldc, bias, residual, ldr, M, N, K, epilogue, out_dtype, norm_weight, norm_eps, stream);
}

extern "C" int sr_gemm_skinny_bulk_bf16(const void* A, int64_t lda, const float* norm_weight, float norm_eps, const void* Wpacked,
                                        void* C, int64_t ldc, const float* bias, const void* residual, int64_t ldr, int M, int N,
                                        int K, int epilogue, int out_dtype, void* stream);
static int g_gemv_bulk = -1;  // -1: env SR_GEMV_BULK (default 0: measured slower than the LDG kernel on B200 except at K >= 8192)
extern "C" int sr_set_gemv_bulk(int on) {
  g_gemv_bulk = on ? 1 : 0;
  return SR_OK;
}

extern "C" int sr_gemm_skinny_packed_bf16(const void* A, int64_t lda, const float* norm_weight, float norm_eps, const void* Wpacked,
                                          void* C, int64_t ldc, const float* bias, const void* residual, int64_t ldr, int M,
                                          int N, int K, int epilogue, int out_dtype, void* stream) {
  if (g_gemv_bulk < 0) {
    const char* e = getenv("SR_GEMV_BULK");
    g_gemv_bulk = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (g_gemv_bulk && M <= 8)  // bulk-copy-engine streaming variant (sr_gemv_bulk.cu)
    return sr_gemm_skinny_bulk_bf16(A, lda, norm_weight, norm_eps, Wpacked, C, ldc, bias, residual, ldr, M, N, K, epilogue,
                                    out_dtype, stream);
  return gemm_skinny_impl(A, lda, Wpacked, K, C, ldc, bias, residual, ldr, M, N, K, epilogue, out_dtype, norm_weight, norm_eps, stream, 1);
}

extern "C" int sr_set_gemv_load(int flavour) {
  if (flavour < 0 || flavour > 2) return SR_ERR_INVALID;
  sr::g_gemv_load = flavour;
  return SR_OK;
}

extern "C" int sr_set_gemv_warps(int warps) {
  if (warps != 0 && warps != 4 && warps != 8 && warps != 16) return SR_ERR_INVALID;
  sr::g_gemv_warps = warps;
  return SR_OK;
}

extern "C" int64_t sr_gemm_skinny_splitk_workspace_bytes(int N) {
  const int64_t groups = (N + 15) / 16;
  return ((groups * 4 + 15) / 16) * 16 + groups * 2 * 128 * 4;
}

extern "C" int sr_gemm_skinny_packed_splitk_bf16(const void* A, int64_t lda, const void* Wpacked, void* C, int64_t ldc,
                                                 const float* bias, const void* residual, int64_t ldr, int M, int N, int K,
                                                 int epilogue, int out_dtype, void* workspace, void* stream) {
  return gemm_skinny_impl(A, lda, Wpacked, K, C, ldc, bias, residual, ldr, M, N, K, epilogue, out_dtype, nullptr, 0.f, stream, 1, 2,
                          workspace);
}
