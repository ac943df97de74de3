// sr_gemv_bulk.cu -- decode-step skinny GEMM (M <= 8) with the weights streamed by the bulk-copy engine.
//
// Same math and epilogues as sr_gemv.cu, different data path: instead of per-thread 16-byte loads (whose outstanding-
// request budget caps a 512-thread CTA at ~3-4.6 TB/s on B200, tools/gemv_bench.py), ONE producer thread issues
// cp.async.bulk copies of 8-16 KB linear slabs of the PACKED weight layout ([N/16][K/32][16][32], 1 KB blocks) into a
// deep shared-memory ring (96-128 KB in flight per SM) guarded by mbarriers; 8 consumer warps read their A fragments from
// shared memory (conflict-free 512-byte warp reads), feed mma.sync.m16n8k16 exactly like sr_gemv.cu (same k permutation,
// so results are bit-identical), and hand slots back.  The producer does not wait on the programmatic dependency --
// weights never depend on the previous kernel -- so the ring is already filling while the previous decode kernel drains.
#include "sr_host.h"

namespace sr {

constexpr int GB_CONSUMERS = 8;                 // consumer warps
constexpr int GB_THREADS = (GB_CONSUMERS + 1) * 32;
constexpr int GB_CH = 8;                        // chunks (1 KB blocks) per row group per stage = one per consumer warp
constexpr int GB_XPF = 3;                       // activation-fragment prefetch distance (stages)

struct GemvBulkParams {
  const __nv_bfloat16* A;
  const __nv_bfloat16* Wp;
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  const float* norm_w;
  int64_t lda, ldc, ldr;
  int M, N, K;
  float norm_eps;
};

SR_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
SR_DEVINL void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int RG>
struct GbCfg {
  static constexpr int kStageBytes = RG * GB_CH * 1024;
  static constexpr int kStages = (RG == 1) ? 12 : 8;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 2 * kStages * 8 + 64;
};

template <int RG, int EPI, bool OUT_F32, bool NORM>
__global__ void __launch_bounds__(GB_THREADS, 1) gemv_bulk_kernel(const GemvBulkParams p) {
  using Cfg = GbCfg<RG>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  __shared__ float red[GB_CONSUMERS][RG][16 * 8];
  __shared__ float ssq_red[GB_CONSUMERS][8];
  __shared__ float inv_rms[8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * (16 * RG);
  const int nchunks = p.K >> 5;
  const int nstages = (nchunks + GB_CH - 1) / GB_CH;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], GB_CONSUMERS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_trigger();

  if (warp == 0) {
    // ---------------- producer: weights only, no dependency on the previous kernel ----------------
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int st = 0; st < nstages; ++st) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        const int nch = min(GB_CH, nchunks - st * GB_CH);
        mbar_arrive_expect_tx(&full_bar[slot], static_cast<uint32_t>(RG * nch * 1024));
#pragma unroll
        for (int r = 0; r < RG; ++r) {
          const __nv_bfloat16* src = p.Wp + (static_cast<int64_t>(row0 / 16 + r) * nchunks + st * GB_CH) * 512;
          bulk_g2s(smem + slot * Cfg::kStageBytes + r * (GB_CH * 1024), src, static_cast<uint32_t>(nch * 1024), &full_bar[slot]);
        }
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int cw = warp - 1;                   // chunk within each stage
  const int g = lane >> 2, t = lane & 3;
  pdl_wait();                                // activations / residual come from the previous kernel
  const bool xok = g < p.M;
  const __nv_bfloat16* xrow = p.A + static_cast<int64_t>(xok ? g : 0) * p.lda + t * 8;
  float acc[RG][4];
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
  float ssq = 0.f;

  uint4 xq[GB_XPF];
  float4 nq0[GB_XPF], nq1[GB_XPF];
  auto load_x = [&](int slot_i, int st) {
    const int c = st * GB_CH + cw;
    if (c < nchunks) {
      xq[slot_i] = xok ? *reinterpret_cast<const uint4*>(xrow + c * 32) : make_uint4(0, 0, 0, 0);
      if (NORM) {
        nq0[slot_i] = __ldg(reinterpret_cast<const float4*>(p.norm_w + c * 32 + t * 8));
        nq1[slot_i] = __ldg(reinterpret_cast<const float4*>(p.norm_w + c * 32 + t * 8 + 4));
      }
    }
  };
#pragma unroll
  for (int i = 0; i < GB_XPF; ++i) load_x(i, i);

  int slot = 0;
  uint32_t phase = 0;
  for (int st0 = 0; st0 < nstages; st0 += GB_XPF) {
#pragma unroll
    for (int i = 0; i < GB_XPF; ++i) {
      const int st = st0 + i;
      if (st < nstages) {
        mbar_wait(&full_bar[slot], phase);
        const int c = st * GB_CH + cw;
        if (c < nchunks) {
          uint4 xv = xq[i];
          if (NORM) {
            const float a0 = bf16_lo(xv.x), a1 = bf16_hi(xv.x), a2 = bf16_lo(xv.y), a3 = bf16_hi(xv.y);
            const float a4 = bf16_lo(xv.z), a5 = bf16_hi(xv.z), a6 = bf16_lo(xv.w), a7 = bf16_hi(xv.w);
            ssq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7;
            xv.x = pack_bf16x2(a0 * nq0[i].x, a1 * nq0[i].y);
            xv.y = pack_bf16x2(a2 * nq0[i].z, a3 * nq0[i].w);
            xv.z = pack_bf16x2(a4 * nq1[i].x, a5 * nq1[i].y);
            xv.w = pack_bf16x2(a6 * nq1[i].z, a7 * nq1[i].w);
          }
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            const uint8_t* blk = smem + slot * Cfg::kStageBytes + r * (GB_CH * 1024) + cw * 1024;
            const uint4 w0 = *reinterpret_cast<const uint4*>(blk + g * 64 + t * 16);
            const uint4 w1 = *reinterpret_cast<const uint4*>(blk + (g + 8) * 64 + t * 16);
            mma16816(acc[r], w0.x, w1.x, w0.y, w1.y, xv.x, xv.y);
            mma16816(acc[r], w0.z, w1.z, w0.w, w1.w, xv.z, xv.w);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
        load_x(i, st + GB_XPF);
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
    }
  }

  // cross-warp reduction + epilogue (consumer warps only: named barrier 1, 256 threads)
#pragma unroll
  for (int r = 0; r < RG; ++r) {
    red[cw][r][g * 8 + 2 * t] = acc[r][0];
    red[cw][r][g * 8 + 2 * t + 1] = acc[r][1];
    red[cw][r][(g + 8) * 8 + 2 * t] = acc[r][2];
    red[cw][r][(g + 8) * 8 + 2 * t + 1] = acc[r][3];
  }
  if (NORM) {
    float v = ssq;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) ssq_red[cw][g] = v;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const int ct = threadIdx.x - 32;           // 0..255
  if (NORM) {
    if (ct < 8) {
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GB_CONSUMERS; ++w8) v += ssq_red[w8][ct];
      inv_rms[ct] = rsqrtf(v / static_cast<float>(p.K) + p.norm_eps);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  if (ct < 128) {
    const int b = ct >> 4, f = ct & 15;      // consecutive threads -> consecutive features
    float v[RG];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GB_CONSUMERS; ++w8) s += red[w8][r][f * 8 + b];
      v[r] = NORM ? s * inv_rms[b] : s;
    }
    if (b < p.M) {
      if (EPI == SR_EPI_SWIGLU) {
        const int prow = row0 + f;           // packed gate row; up row = prow + 16
        float gate = v[0], up = v[RG - 1];
        if (p.bias) { gate += __ldg(p.bias + prow); up += __ldg(p.bias + prow + 16); }
        const float o = __fdividef(gate, 1.0f + __expf(-gate)) * up;
        reinterpret_cast<__nv_bfloat16*>(p.C)[static_cast<int64_t>(b) * p.ldc + (row0 >> 1) + f] = __float2bfloat16(o);
      } else {
        const int col = row0 + f;
        float o = v[0];
        if (p.bias) o += __ldg(p.bias + col);
        if (EPI == SR_EPI_GELU) o = 0.5f * o * (1.0f + erff(o * 0.70710678118654752440f));
        if (EPI == SR_EPI_RESIDUAL) o += __bfloat162float(p.residual[static_cast<int64_t>(b) * p.ldr + col]);
        if (OUT_F32) reinterpret_cast<float*>(p.C)[static_cast<int64_t>(b) * p.ldc + col] = o;
        else reinterpret_cast<__nv_bfloat16*>(p.C)[static_cast<int64_t>(b) * p.ldc + col] = __float2bfloat16(o);
      }
    }
  }
}

template <int RG, int EPI, bool OUT_F32, bool NORM>
static int launch_bulk(const GemvBulkParams& p, int grid, cudaStream_t s) {
  using Cfg = GbCfg<RG>;
  auto kern = gemv_bulk_kernel<RG, EPI, OUT_F32, NORM>;
  static unsigned long long attr_done = 0;  // per device
  SR_CHECK_CUDA(ensure_dynamic_smem(kern, Cfg::kSmemBytes, &attr_done));
  SR_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GB_THREADS), static_cast<size_t>(Cfg::kSmemBytes), s, p));
  return SR_OK;
}

template <bool NORM>
static int dispatch_bulk(const GemvBulkParams& p, int epi, int out_f32, cudaStream_t s) {
  if (epi == SR_EPI_SWIGLU) return launch_bulk<2, SR_EPI_SWIGLU, false, NORM>(p, p.N / 32, s);
  const int grid = p.N / 16;
  if (out_f32) return launch_bulk<1, SR_EPI_NONE, true, NORM>(p, grid, s);
  if (epi == SR_EPI_NONE) return launch_bulk<1, SR_EPI_NONE, false, NORM>(p, grid, s);
  if (epi == SR_EPI_RESIDUAL) return launch_bulk<1, SR_EPI_RESIDUAL, false, NORM>(p, grid, s);
  if (epi == SR_EPI_GELU) return launch_bulk<1, SR_EPI_GELU, false, NORM>(p, grid, s);
  set_last_error("sr_gemm_skinny_bulk_bf16: unknown epilogue %d", epi);
  return SR_ERR_INVALID;
}

}  // namespace sr

extern "C" int sr_gemm_skinny_bulk_bf16(const void* A, int64_t lda, const float* norm_weight, float norm_eps, const void* Wpacked,
                                        void* C, int64_t ldc, const float* bias, const void* residual, int64_t ldr, int M, int N,
                                        int K, int epilogue, int out_dtype, void* stream) {
  using namespace sr;
  SR_REQUIRE(A && Wpacked && C, "sr_gemm_skinny_bulk_bf16: null operand");
  SR_REQUIRE(M > 0 && M <= 8 && N > 0 && N % 16 == 0 && K > 0 && K % 32 == 0, "sr_gemm_skinny_bulk_bf16: need 0 < M <= 8, N %% 16 == 0, K %% 32 == 0");
  SR_REQUIRE(lda % 8 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wpacked) & 15) == 0,
             "sr_gemm_skinny_bulk_bf16: alignment");
  SR_REQUIRE(epilogue != SR_EPI_RESIDUAL || residual, "sr_gemm_skinny_bulk_bf16: residual epilogue needs residual");
  SR_REQUIRE(epilogue != SR_EPI_SWIGLU || (N % 32 == 0 && out_dtype == SR_DTYPE_BF16), "sr_gemm_skinny_bulk_bf16: SwiGLU needs N %% 32 == 0, bf16 out");
  SR_REQUIRE(out_dtype == SR_DTYPE_BF16 || epilogue == SR_EPI_NONE, "sr_gemm_skinny_bulk_bf16: fp32 output supports SR_EPI_NONE only");
  SR_REQUIRE(norm_weight == nullptr || (reinterpret_cast<uintptr_t>(norm_weight) & 15) == 0, "sr_gemm_skinny_bulk_bf16: norm_weight alignment");
  GemvBulkParams p;
  p.A = reinterpret_cast<const __nv_bfloat16*>(A);
  p.Wp = reinterpret_cast<const __nv_bfloat16*>(Wpacked);
  p.C = C;
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.norm_w = norm_weight;
  p.lda = lda; p.ldc = ldc; p.ldr = ldr;
  p.M = M; p.N = N; p.K = K;
  p.norm_eps = norm_eps;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int f32 = out_dtype == SR_DTYPE_F32;
  return norm_weight ? dispatch_bulk<true>(p, epilogue, f32, s) : dispatch_bulk<false>(p, epilogue, f32, s);
}
