// sr_gemm.cu -- C[M,N] = epilogue(A[M,K] . W[N,K]^T) on tcgen05 tensor cores (sm_100a).
//
// Replaces every nn.Linear the reference reaches through cuBLAS on the hot path (SURVEY.md 2.2 rows
// K1,K4,K8,K9,K10,K13,K15,K16): HF modeling_qwen2_5_vl.py:91-114 (patch embed as GEMM), :214/:285
// (ViT qkv/proj), :77-88 (ViT SwiGLU MLP), :133-146 (merger), :704-707 (decoder q/k/v/o), :611-624
// (decoder MLP), :1385 (lm_head).
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2D boxes [64 x 128] (A) and [64 x BLOCK_N] (W), 128B swizzle,
//              kStages-deep mbarrier ring.
//   warp 1   : MMA issuer    -- one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x4 per k-block into a
//              double-buffered TMEM accumulator (2 x BLOCK_N fp32 columns); tcgen05.commit frees smem stages
//              and publishes finished accumulators.
//   warps 2-9: epilogue      -- tcgen05.ld 32 lanes x 32 columns, fused bias / residual / SwiGLU / GELU,
//              bf16 (or fp32) vector stores.  Overlaps with the next tile's MMAs through the 2nd TMEM buffer.
//              EIGHT warps (two per TMEM lane quarter, each taking half of the tile's columns): with four, one
//              latency-exposed warp per scheduler needed ~2100 cycles per 32-column chunk = 16 900 cycles per
//              128 x 256 tile, more than the 10 240 MMA cycles of a K = 1280 tile -- the short-K shapes were
//              epilogue-paced (profiles/r01_gemm_vit_qkv.json: 320 865 cycles for 19 tile waves).  The residual rows of
//              a warp's whole column half are requested before it waits for the accumulator.
// Both operands are K-major (activations [M,K] row-major; weights in nn.Linear layout [N,K]).
#include "sr_host.h"
#include "../../include/socioreasoner_b200.h"

namespace sr {

constexpr int SR_EPI_LSE = 4;  // internal: no C; per (row, N tile) softmax statistics for sr_lmhead_logprob_entropy
constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
#ifndef SR_GEMM_EPI_WARPS
#define SR_GEMM_EPI_WARPS 8  // 4 or 8 (compile-time A/B switch; 8 is the product configuration)
#endif
// epilogue warps per instantiation: the LSE epilogue keeps one warp per row quarter (its per-row statistics run over all
// columns of the tile), as do the 32-column tiles (a single chunk)
template <int BLOCK_N, int EPI>
struct GemmEpi {
  static constexpr int kWarps = (EPI == SR_EPI_LSE || BLOCK_N < 64) ? 4 : SR_GEMM_EPI_WARPS;
};
constexpr int GEMM_MAX_THREADS = 64 + 32 * SR_GEMM_EPI_WARPS;

struct GemmParams {
  void* C;
  const float* bias;      // [N] fp32 or null
  const void* residual;   // bf16 [M, ldr] or null
  int64_t ldc, ldr;
  int M, N, K;
  int num_m_blocks, num_n_blocks;
  int band_rows;          // tile order: row blocks per band (gemm_tile_coords)
  int wide;               // 1: C (and the residual) rows allow 32-byte accesses -> 256-bit LDG / STG (one full sector per thread)
  const int64_t* labels;  // SR_EPI_LSE: [M] target column per row (out of range: no target)
  float4* partial;        // SR_EPI_LSE: [num_n_blocks][M] (max, sum exp(x - max), sum exp(x - max) * x, x[label] or 0)
};

// Tile order: bands of GEMM_BAND row blocks, inside a band the row block runs fastest and the column block slowest.  The ~148
// tiles in flight then cover (a few column blocks) x (one band): the band's A rows (GEMM_BAND x 128 x K) are fetched from HBM
// once and re-used from L2 by every column block, and neighbouring CTAs share B tiles.  Row-major-over-M order (round 1) swept
// all of A once per column block: 1.46 x the algorithmic DRAM bytes on the ViT qkv shape and 5 x A for the N = 1280 shapes
// whose A (161 MB at K = 3456) does not fit in L2.
// Band height (chosen on the host, measured with tools/gemm_bench.py): 37 row blocks (a wave = 4 column blocks x 37 row blocks:
// weights larger than L2, e.g. the decoder gate/up at 90 MB, are then swept only once per 37 row blocks), 8 for long-K shapes
// with small weights (ViT fc2, decoder down: A is the large operand).
SR_DEVINL void gemm_tile_coords(int tile, int num_m_blocks, int num_n_blocks, int band_rows, int& m_blk, int& n_blk) {
  const int per_band = band_rows * num_n_blocks;
  const int band = tile / per_band;
  const int r = tile - band * per_band;
  const int m0 = band * band_rows;
  const int rows = min(band_rows, num_m_blocks - m0);
  n_blk = r / rows;
  m_blk = m0 + (r - n_blk * rows);
}

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 128) ? 6 : 8);
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
  static constexpr int kBiasBytes = 2 * BLOCK_N * 4;  // double-buffered per-tile bias slice (epilogue warps)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes;
};

// 256-bit global accesses (sm_100: LDG.256 / STG.256).  A thread of the epilogue owns one output ROW, so its stores can never
// coalesce with its neighbours'; 32-byte stores at least write whole sectors (16-byte ones leave L2 with two partial-sector
// writes per sector, twice the requests).
SR_DEVINL void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w)
               : "memory");
}
SR_DEVINL void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
SR_DEVINL void store_row16(void* dst, const uint4& a, const uint4& b, bool wide) {
  if (wide) {
    stg256(dst, a, b);
  } else {
    *reinterpret_cast<uint4*>(dst) = a;
    *(reinterpret_cast<uint4*>(dst) + 1) = b;
  }
}

SR_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
SR_DEVINL float silu(float x) { return x / (1.0f + __expf(-x)); }

template <int BLOCK_N, int EPI, bool OUT_F32>
__global__ void __launch_bounds__(GEMM_MAX_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kEpiWarps = GemmEpi<BLOCK_N, EPI>::kWarps;
  constexpr int kEpiThreads = 32 * kEpiWarps;
  constexpr int kChunks = BLOCK_N / 32 / (kEpiWarps / 4);  // 32-column chunks per epilogue warp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256);  // [2][BLOCK_N]

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], kEpiWarps);
    mbar_init(&tmem_empty_bar[1], kEpiWarps);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_k_blocks = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  if (warp_idx == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        gemm_tile_coords(tile, p.num_m_blocks, p.num_n_blocks, p.band_rows, m_blk, n_blk);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BLOCK_K, m_blk * GEMM_BLOCK_M);
          tma_load_2d(sa + Cfg::kABytes, &tmB, &full_bar[stage], kb * GEMM_BLOCK_K, n_blk * BLOCK_N);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16_ss(tmem_d, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------- epilogue warps (2 .. 2 + kEpiWarps) ----------------
    // Global-memory operands of the epilogue are kept off the critical path: the tile's bias slice is staged once in
    // shared memory (a per-use __ldg costs one L2 round trip per 8 columns -- ncu showed 60% of all stall samples
    // there), and the residual rows of the warp's column half are requested before the accumulator is waited for.
    const int quarter = warp_idx & 3;           // TMEM lane quarter this warp may access
    const int half = (warp_idx - 2) >> 2;       // which kChunks-chunk column range of the tile
    const int et = threadIdx.x - 64;            // 0 .. kEpiThreads-1 within the epilogue group
    constexpr int kResChunks = (EPI == SR_EPI_RESIDUAL) ? kChunks : 1;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tile_par = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      gemm_tile_coords(tile, p.num_m_blocks, p.num_n_blocks, p.band_rows, m_blk, n_blk);
      float* sb = s_bias + tile_par * BLOCK_N;
      if (p.bias != nullptr) {
        for (int j = et; j < BLOCK_N; j += kEpiThreads) {
          const int col = n_blk * BLOCK_N + j;
          sb[j] = col < p.N ? __ldg(p.bias + col) : 0.f;
        }
      }
      const int row = m_blk * GEMM_BLOCK_M + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int chunk0 = half * kChunks;
      uint4 rres[kResChunks][4];
      if (EPI == SR_EPI_RESIDUAL) {
        const __nv_bfloat16* rrow = reinterpret_cast<const __nv_bfloat16*>(p.residual) + static_cast<int64_t>(row_ok ? row : 0) * p.ldr;
#pragma unroll
        for (int c = 0; c < kResChunks; ++c) {
#pragma unroll
          for (int g16 = 0; g16 < 2; ++g16) {
            const int col = n_blk * BLOCK_N + (chunk0 + c) * 32 + g16 * 16;
            rres[c][2 * g16] = rres[c][2 * g16 + 1] = make_uint4(0, 0, 0, 0);
            if (row_ok && col < p.N) {
              if (p.wide) {
                ldg256(rrow + col, rres[c][2 * g16], rres[c][2 * g16 + 1]);
              } else {
                rres[c][2 * g16] = *reinterpret_cast<const uint4*>(rrow + col);
                if (col + 8 < p.N) rres[c][2 * g16 + 1] = *reinterpret_cast<const uint4*>(rrow + col + 8);
              }
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // bias slice visible to the epilogue warps
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
      float lse_m = -INFINITY, lse_s = 0.f, lse_w = 0.f, lse_t = 0.f;
      int lse_label = -1;
      if (EPI == SR_EPI_LSE && row_ok) {
        const int64_t lb = p.labels[row];
        lse_label = (lb >= 0 && lb < p.N) ? static_cast<int>(lb) : -1;
      }
#pragma unroll
      for (int ci = 0; ci < kChunks; ++ci) {
        const int chunk = chunk0 + ci;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(chunk * 32), v);
        const int col0 = n_blk * BLOCK_N + chunk * 32;
        if (EPI == SR_EPI_LSE) {
          // online softmax statistics over the tile's columns (the logits never leave the SM):
          // log Z = m + log s,  sum p x = w / s,  x[label]  (roll/utils/functionals.py:193-205)
          tmem_ld_wait();
          if (col0 < p.N) {
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = fmaxf(cm, (col0 + j < p.N) ? __uint_as_float(v[j]) : -INFINITY);
            const float mn = fmaxf(lse_m, cm);
            const float sc = __expf(lse_m - mn);  // exp(-inf) = 0 on the first chunk
            lse_s *= sc;
            lse_w *= sc;
            lse_m = mn;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = __uint_as_float(v[j]);
              const float e = (col0 + j < p.N) ? __expf(x - mn) : 0.f;
              lse_s += e;
              lse_w = fmaf(e, (col0 + j < p.N) ? x : 0.f, lse_w);
              if (col0 + j == lse_label) lse_t = x;
            }
          }
          continue;
        }
        tmem_ld_wait();
        const float* sbc = sb + chunk * 32;
        if (EPI == SR_EPI_SWIGLU) {
          // packed weight rows: [16 gate | 16 up] per 32 columns -> 16 outputs at col0/2
          const int oc0 = col0 >> 1;
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float g = __uint_as_float(v[j]);
            float u = __uint_as_float(v[16 + j]);
            if (p.bias != nullptr) {
              g += sbc[j];
              u += sbc[16 + j];
            }
            o[j] = __fdividef(g, 1.0f + __expf(-g)) * u;
          }
          if (row_ok && oc0 < (p.N >> 1)) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc + oc0;
            uint4 pk[2];
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
              pk[g8].x = pack_bf16x2(o[g8 * 8 + 0], o[g8 * 8 + 1]);
              pk[g8].y = pack_bf16x2(o[g8 * 8 + 2], o[g8 * 8 + 3]);
              pk[g8].z = pack_bf16x2(o[g8 * 8 + 4], o[g8 * 8 + 5]);
              pk[g8].w = pack_bf16x2(o[g8 * 8 + 6], o[g8 * 8 + 7]);
            }
            store_row16(dst, pk[0], pk[1], p.wide != 0);
          }
        } else {
#pragma unroll
          for (int g16 = 0; g16 < 2; ++g16) {
            const int col = col0 + g16 * 16;
            if (col >= p.N) break;
            const bool second = col + 8 < p.N;  // N % 8 == 0 (checked on the host); N % 16 == 0 whenever p.wide
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[g16 * 16 + j]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) x[j] += sbc[g16 * 16 + j];
            }
            if (EPI == SR_EPI_GELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) x[j] = gelu_erf(x[j]);
            }
            if (row_ok) {
              if (EPI == SR_EPI_RESIDUAL) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint4 r = rres[EPI == SR_EPI_RESIDUAL ? ci : 0][2 * g16 + h];
                  float* xx = x + 8 * h;
                  xx[0] += bf16_lo(r.x); xx[1] += bf16_hi(r.x); xx[2] += bf16_lo(r.y); xx[3] += bf16_hi(r.y);
                  xx[4] += bf16_lo(r.z); xx[5] += bf16_hi(r.z); xx[6] += bf16_lo(r.w); xx[7] += bf16_hi(r.w);
                }
              }
              if (OUT_F32) {
                float* dst = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  if (h == 1 && !second) break;
                  const float* xx = x + 8 * h;
                  const uint4 a = make_uint4(__float_as_uint(xx[0]), __float_as_uint(xx[1]), __float_as_uint(xx[2]), __float_as_uint(xx[3]));
                  const uint4 b = make_uint4(__float_as_uint(xx[4]), __float_as_uint(xx[5]), __float_as_uint(xx[6]), __float_as_uint(xx[7]));
                  store_row16(dst + 8 * h, a, b, p.wide != 0);
                }
              } else {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc + col;
                uint4 pk[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  pk[h].x = pack_bf16x2(x[8 * h + 0], x[8 * h + 1]);
                  pk[h].y = pack_bf16x2(x[8 * h + 2], x[8 * h + 3]);
                  pk[h].z = pack_bf16x2(x[8 * h + 4], x[8 * h + 5]);
                  pk[h].w = pack_bf16x2(x[8 * h + 6], x[8 * h + 7]);
                }
                if (p.wide) {
                  stg256(dst, pk[0], pk[1]);
                } else {
                  *reinterpret_cast<uint4*>(dst) = pk[0];
                  if (second) *reinterpret_cast<uint4*>(dst + 8) = pk[1];
                }
              }
            }
          }
        }
      }
      if (EPI == SR_EPI_LSE && row_ok)
        p.partial[static_cast<int64_t>(n_blk) * p.M + row] = make_float4(lse_m, lse_s, lse_w, lse_t);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      tile_par ^= 1;
    }
  }

  // ---------------- teardown ----------------
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, int EPI, bool OUT_F32>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, EPI, OUT_F32>;
  static unsigned long long attr_done = 0;  // per device
  SR_CHECK_CUDA(ensure_dynamic_smem(kern, Cfg::kSmemBytes, &attr_done));
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  kern<<<grid, 64 + 32 * GemmEpi<BLOCK_N, EPI>::kWarps, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

template <int BLOCK_N>
static int dispatch_epi(int epi, int out_f32, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s) {
  if (out_f32) {
    if (epi != SR_EPI_NONE) {
      set_last_error("fp32 output supports SR_EPI_NONE only");
      return SR_ERR_UNSUPPORTED;
    }
    return launch_gemm<BLOCK_N, SR_EPI_NONE, true>(a, b, p, s);
  }
  switch (epi) {
    case SR_EPI_NONE: return launch_gemm<BLOCK_N, SR_EPI_NONE, false>(a, b, p, s);
    case SR_EPI_RESIDUAL: return launch_gemm<BLOCK_N, SR_EPI_RESIDUAL, false>(a, b, p, s);
    case SR_EPI_SWIGLU: return launch_gemm<BLOCK_N, SR_EPI_SWIGLU, false>(a, b, p, s);
    case SR_EPI_GELU: return launch_gemm<BLOCK_N, SR_EPI_GELU, false>(a, b, p, s);
  }
  set_last_error("unknown epilogue %d", epi);
  return SR_ERR_INVALID;
}

// Merge of the per-tile statistics: one warp per row, lanes stride over the N tiles, online merge through shuffles.
__global__ void lse_merge_kernel(const float4* __restrict__ partial, int M, int n_tiles, float* __restrict__ logprob,
                                 float* __restrict__ entropy) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float m = -INFINITY, s = 0.f, w = 0.f, t = 0.f;
  for (int nb = lane; nb < n_tiles; nb += 32) {
    const float4 q = __ldg(partial + static_cast<int64_t>(nb) * M + row);
    const float mn = fmaxf(m, q.x);
    const float a = __expf(m - mn), b = __expf(q.x - mn);
    s = s * a + q.y * b;
    w = w * a + q.z * b;
    t += q.w;
    m = mn;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, d), s2 = __shfl_xor_sync(0xffffffffu, s, d);
    const float w2 = __shfl_xor_sync(0xffffffffu, w, d), t2 = __shfl_xor_sync(0xffffffffu, t, d);
    const float mn = fmaxf(m, m2);
    const float a = (m == -INFINITY) ? 0.f : __expf(m - mn), b = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
    s = s * a + s2 * b;
    w = w * a + w2 * b;
    t += t2;
    m = mn;
  }
  if (lane == 0) {
    const float logz = m + logf(s);
    if (logprob != nullptr) logprob[row] = t - logz;
    if (entropy != nullptr) entropy[row] = logz - w / s;
  }
}

}  // namespace sr

extern "C" int sr_gemm_skinny_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                                   const float* bias, const void* residual, int64_t ldr, int M, int N, int K, int epilogue,
                                   int out_dtype, void* stream);

extern "C" int sr_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                            const float* bias, const void* residual, int64_t ldr, int M, int N, int K, int epilogue,
                            int out_dtype, int block_n, void* stream) {
  using namespace sr;
  SR_REQUIRE(A && W && C, "sr_gemm_bf16: null operand");
  SR_REQUIRE(M > 0 && N > 0 && K > 0, "sr_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  SR_REQUIRE(N % 8 == 0, "sr_gemm_bf16: N=%d must be a multiple of 8", N);
  SR_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "sr_gemm_bf16: lda/ldw must be >=K and multiples of 8");
  SR_REQUIRE(epilogue != SR_EPI_RESIDUAL || (residual && ldr % 8 == 0), "sr_gemm_bf16: residual epilogue needs residual, ldr%%8==0");
  const int n_out = (epilogue == SR_EPI_SWIGLU) ? N / 2 : N;
  SR_REQUIRE(epilogue != SR_EPI_SWIGLU || N % 32 == 0, "sr_gemm_bf16: SwiGLU needs N %% 32 == 0 (16-gate/16-up packing)");
  SR_REQUIRE(ldc >= n_out && ldc % (out_dtype == SR_DTYPE_F32 ? 4 : 8) == 0, "sr_gemm_bf16: bad ldc=%lld", (long long)ldc);
  SR_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0, "sr_gemm_bf16: C not 16B aligned");
  SR_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "sr_gemm_bf16: bias not 16B aligned");

  if (block_n == 0 && M <= 32 && K % 32 == 0)  // decode-step shapes: weight-streaming kernel (sr_gemv.cu), HBM-bound
    return sr_gemm_skinny_bf16(A, lda, W, ldw, C, ldc, bias, residual, ldr, M, N, K, epilogue, out_dtype, stream);
  int bn = block_n;
  if (bn == 0) bn = (M <= 64) ? 32 : 256;
  SR_REQUIRE(bn == 32 || bn == 128 || bn == 256, "sr_gemm_bf16: block_n must be 0/32/128/256");

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)GEMM_BLOCK_K, (uint32_t)GEMM_BLOCK_M};
    int rc = encode_tmap_bf16(&tmA, A, 2, dims, strides, box, 128);
    if (rc != SR_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)GEMM_BLOCK_K, (uint32_t)bn};
    int rc = encode_tmap_bf16(&tmB, W, 2, dims, strides, box, 128);
    if (rc != SR_OK) return rc;
  }
  GemmParams p = {};
  p.C = C;
  p.bias = bias;
  p.residual = residual;
  p.ldc = ldc;
  p.ldr = ldr;
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  p.num_n_blocks = (N + bn - 1) / bn;
  const int f32 = out_dtype == SR_DTYPE_F32;
  p.band_rows = (K >= 3072 && static_cast<int64_t>(N) * K * 2 <= (64ll << 20)) ? 8 : 37;
  {  // 32-byte accesses: every row start and every 16-column group (8 for fp32 / SwiGLU halves) must be 32-byte aligned
    const int64_t c_el = f32 ? 4 : 2;
    bool ok = (reinterpret_cast<uintptr_t>(C) & 31) == 0 && (ldc * c_el) % 32 == 0 && N % 16 == 0;
    if (epilogue == SR_EPI_SWIGLU) ok = ok && N % 32 == 0;
    if (epilogue == SR_EPI_RESIDUAL) ok = ok && (reinterpret_cast<uintptr_t>(residual) & 31) == 0 && (ldr * 2) % 32 == 0;
    p.wide = ok ? 1 : 0;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (bn == 256) return dispatch_epi<256>(epilogue, f32, tmA, tmB, p, s);
  if (bn == 128) return dispatch_epi<128>(epilogue, f32, tmA, tmB, p, s);
  return dispatch_epi<32>(epilogue, f32, tmA, tmB, p, s);
}


// ---------------------------------------------------------------------------------------------------------------------
// Fused LM head + log_probs_from_logits / entropy_from_logits: logits = A . W^T are reduced tile by tile inside the GEMM
// epilogue; the [M, V] fp32 logits (0.6 MB per position for V = 151936) are never written.
extern "C" int64_t sr_lmhead_logprob_workspace_bytes(int M, int N) {
  const int64_t nb = (N + 255) / 256;
  return nb * static_cast<int64_t>(M) * 16;
}

extern "C" int sr_lmhead_logprob_entropy(const void* A, int64_t lda, const void* W, int64_t ldw, const int64_t* labels, int M, int N,
                                         int K, void* workspace, float* logprob, float* entropy, void* stream) {
  using namespace sr;
  SR_REQUIRE(A && W && labels && workspace, "sr_lmhead_logprob_entropy: null operand");
  SR_REQUIRE(logprob || entropy, "sr_lmhead_logprob_entropy: no output requested");
  SR_REQUIRE(M > 0 && N > 0 && K > 0, "sr_lmhead_logprob_entropy: bad shape M=%d N=%d K=%d", M, N, K);
  SR_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "sr_lmhead_logprob_entropy: lda/ldw must be >=K and multiples of 8");
  SR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sr_lmhead_logprob_entropy: workspace not 16B aligned");
  constexpr int bn = 256;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)GEMM_BLOCK_K, (uint32_t)GEMM_BLOCK_M};
    int rc = encode_tmap_bf16(&tmA, A, 2, dims, strides, box, 128);
    if (rc != SR_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)GEMM_BLOCK_K, (uint32_t)bn};
    int rc = encode_tmap_bf16(&tmB, W, 2, dims, strides, box, 128);
    if (rc != SR_OK) return rc;
  }
  GemmParams p = {};
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  p.num_n_blocks = (N + bn - 1) / bn;
  p.labels = labels;
  p.band_rows = 37;
  p.partial = reinterpret_cast<float4*>(workspace);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc = launch_gemm<bn, SR_EPI_LSE, false>(tmA, tmB, p, s);
  if (rc != SR_OK) return rc;
  constexpr int warps = 8;
  lse_merge_kernel<<<(M + warps - 1) / warps, warps * 32, 0, s>>>(p.partial, M, p.num_n_blocks, logprob, entropy);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}
