// sr_attention.cu -- variable-length softmax(Q K^T * scale) V on tcgen05 tensor cores (sm_100a).
//
// One kernel serves the three attention shapes of the hot path:
//   * ViT windowed attention  (HF modeling_qwen2_5_vl.py:244-283 with cu_window_seqlens; 28 of 32 blocks)
//   * ViT full attention      (same code with per-image cu_seqlens; blocks 7/15/23/31)
//   * decoder causal GQA prefill (HF:740-756; 16 q heads / 2 kv heads, head_dim 128)
// which the reference reaches through flash-attn 2 (mma.sync) or an eager per-window python loop.
//
// Generic formulation: query row r attends keys k with row_lo[r] <= k < row_hi[r] (absolute key-row indices).
// A work tile = up to 128 consecutive query rows + the key range [kv_lo, kv_hi) that covers all their ranges
// (host-built list, see socioreasoner_b200/attention_plan.py).  Windows, images, samples and causality are all
// expressed through (tiles, row_lo, row_hi); rows of one tile never straddle samples.
//
// Layouts (bf16):  Q [rows, heads, 128]  K [rows, kv_heads, 128]  (head_dim 80 of the ViT zero-padded to 128)
//                  Vt [kv_heads * HDV, key columns]  (V transposed, so both MMAs are K-major)   O [rows, heads, HDV]
//
// CTA = (tile, q head).  warp 0: TMA producer (Q once, K/V 2-stage ring).  warp 1: one thread issues
// S = Q K^T (M128 N128 K128) into a double-buffered TMEM S and PV = P V (M128 N=HDV K128) into TMEM.
// warps 2-5: one query row per thread: tcgen05.ld S, mask, online softmax in fp32 (exp2), P -> bf16 into
// 128B-swizzled smem (A operand of the PV MMA), O accumulated in registers with the usual rescale.
//
// Two instantiations per head size.  STAGES = 2: one CTA per SM, K/V 2-stage ring, S double-buffered in TMEM (512 columns),
// the next block's S = Q K^T overlaps this block's softmax.  STAGES = 1 ("lean"): TWO CTAs per SM -- one K/V stage, one S
// buffer (256 TMEM columns), and the P tile ALIASES the K tile (K is dead once S has been committed), 84-96 KB of shared
// memory -- so that the second CTA's loads / MMAs / softmax fill the bubbles of the first one's dependency chain
// (load -> S -> softmax -> P V -> rescale).  The per-block chain is latency-bound (ncu: tensor pipe 6-14 % busy at one CTA
// per SM), which is why the second resident CTA pays off, most of all for the single-block windowed tiles.
#include <stdlib.h>

#include "sr_host.h"

namespace sr {

constexpr int ATT_THREADS = 192;
constexpr int ATT_BM = 128;   // query rows per tile
constexpr int ATT_BN = 128;   // keys per block
constexpr int ATT_D = 128;    // (padded) qk head dim

SR_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  __nv_bfloat16* O;
  int64_t o_row_stride, o_head_stride;
  const int4* tiles;  // {q_row0, q_nrows, kv_lo, kv_hi}
  const int* row_lo;
  const int* row_hi;
  float scale_log2;
  int group;  // q heads per kv head
  int n_tiles, n_heads;  // work items of the tensor-memory variant: (tile, head), head fastest
  long long* trace;      // optional (sr_set_attention_trace): clock64 stamps of CTAs 0..3, [cta][item < 16][16 slots]
};
SR_DEVINL void att_stamp(const AttnParams& p, int it, int slot) {
  if (p.trace != nullptr && blockIdx.x < 4 && it < 16) p.trace[(blockIdx.x * 16 + it) * 16 + slot] = clock64();
}

template <int HDV, int STAGES>
struct AttnCfg {
  static constexpr bool kLean = STAGES == 1;
  static constexpr int kQBytes = ATT_BM * ATT_D * 2;      // 32 KB (2 boxes)
  static constexpr int kKBytes = ATT_BN * ATT_D * 2;      // 32 KB (2 boxes)
  static constexpr int kVBoxBytes = HDV * 128;            // [HDV rows x 64 keys]
  static constexpr int kVBytes = 2 * kVBoxBytes;
  static constexpr int kKVStage = kKBytes + ((kVBytes + 1023) / 1024) * 1024;
  static constexpr int kPBytes = ATT_BM * ATT_BN * 2;     // 32 KB (lean: aliases the K tile of the single stage)
  static constexpr int kSmemBytes = kQBytes + STAGES * kKVStage + (kLean ? 0 : kPBytes) + 1024 + 256;
  static constexpr int kTmemCols = kLean ? 256 : 512;     // S (x STAGES) + PV
  static constexpr uint32_t kTmemPV = kLean ? 128 : 256;
};

template <int HDV, int STAGES>
__global__ void __launch_bounds__(ATT_THREADS, STAGES == 1 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Cfg = AttnCfg<HDV, STAGES>;
  constexpr bool LEAN = Cfg::kLean;
  static_assert(HDV % 16 == 0 && HDV <= 128 && (STAGES == 1 || STAGES == 2), "HDV / STAGES");
  // stage and mbarrier parity of KV block j
  auto st_of = [](int j) { return LEAN ? 0 : (j & 1); };
  auto ph_of = [](int j) { return static_cast<uint32_t>(LEAN ? (j & 1) : ((j >> 1) & 1)); };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::kQBytes;
  uint8_t* sP = LEAN ? sKV : sKV + 2 * Cfg::kKVStage;  // lean: P overwrites K after S = Q K^T has been committed
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * Cfg::kKVStage + (LEAN ? 0 : Cfg::kPBytes));
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_ready = bars + 7;
  uint64_t* pv_full = bars + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int4 tile = p.tiles[blockIdx.x];
  // TMA needs 16-byte-aligned box origins: the key axis is the INNER dimension of the V^T map, so the key range is
  // widened down to a multiple of 8 keys (the extra keys are masked by row_lo like any other out-of-range key).
  const int q_row0 = tile.x, q_nrows = tile.y, kv_lo = tile.z & ~7, kv_hi = tile.w;
  const int head = blockIdx.y;
  const int kv_head = head / p.group;
  const int nblk = (kv_hi - kv_lo + ATT_BN - 1) / ATT_BN;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_ready, 128);
    mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kTmemPV = Cfg::kTmemPV;

  if (nblk > 0) {
    if (warp_idx == 0) {
      // ---------------- TMA producer ----------------
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, Cfg::kQBytes);
        tma_load_3d(sQ, &tmQ, q_full, 0, head, q_row0);
        tma_load_3d(sQ + 16384, &tmQ, q_full, 64, head, q_row0);
        for (int j = 0; j < nblk; ++j) {
          const int s = st_of(j);
          const uint32_t ph = ph_of(j);
          mbar_wait(&kv_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], Cfg::kKBytes + Cfg::kVBytes);
          uint8_t* sK = sKV + s * Cfg::kKVStage;
          uint8_t* sV = sK + Cfg::kKBytes;
          const int kbase = kv_lo + j * ATT_BN;
          tma_load_3d(sK, &tmK, &kv_full[s], 0, kv_head, kbase);
          tma_load_3d(sK + 16384, &tmK, &kv_full[s], 64, kv_head, kbase);
          tma_load_2d(sV, &tmV, &kv_full[s], kbase, kv_head * HDV);
          tma_load_2d(sV + Cfg::kVBoxBytes, &tmV, &kv_full[s], kbase + 64, kv_head * HDV);
        }
      }
    } else if (warp_idx == 1) {
      // ---------------- MMA issuer ----------------
      if (lane == 0) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, HDV);
        const uint32_t aQ = smem_u32(sQ);
        const uint32_t aP = smem_u32(sP);
        auto issue_s = [&](int j) {
          const int s = st_of(j);
          mbar_wait(&kv_full[s], ph_of(j));
          tc_fence_after();
          const uint32_t aK = smem_u32(sKV + s * Cfg::kKVStage);
          const uint32_t tS = tmem_base + static_cast<uint32_t>(s * ATT_BN);
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(aQ + (k >> 2) * 16384) + static_cast<uint64_t>((k & 3) * 2);
            const uint64_t db = umma_desc_k_sw128(aK + (k >> 2) * 16384) + static_cast<uint64_t>((k & 3) * 2);
            umma_bf16_ss(tS, da, db, idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[s]);
        };
        mbar_wait(q_full, 0);
        if (!LEAN) issue_s(0);
        for (int j = 0; j < nblk; ++j) {
          // lean: one S buffer and one K/V stage -- S_j is issued here, after P V of block j-1 (program order on the tensor
          // pipe; the softmax warps finished reading S_{j-1} before they signalled p_ready for block j-1)
          if (LEAN) issue_s(j);
          else if (j + 1 < nblk) issue_s(j + 1);
          mbar_wait(p_ready, j & 1);
          tc_fence_after();
          const int s = st_of(j);
          const uint32_t aV = smem_u32(sKV + s * Cfg::kKVStage + Cfg::kKBytes);
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(aP + (k >> 2) * 16384) + static_cast<uint64_t>((k & 3) * 2);
            const uint64_t db = umma_desc_k_sw128(aV + (k >> 2) * Cfg::kVBoxBytes) + static_cast<uint64_t>((k & 3) * 2);
            umma_bf16_ss(tmem_base + kTmemPV, da, db, idesc_pv, k != 0 ? 1u : 0u);
          }
          umma_commit(pv_full);
          umma_commit(&kv_empty[s]);
        }
      }
    } else {
      // ---------------- softmax / output warps: one query row per thread ----------------
      const int quarter = warp_idx & 3;
      const int r = quarter * 32 + lane;
      const bool row_ok = r < q_nrows;
      int lo = 0, hi = 0;
      if (row_ok) {
        lo = __ldg(p.row_lo + q_row0 + r);
        hi = __ldg(p.row_hi + q_row0 + r);
      }
      const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
      float m = -INFINITY, l = 0.f;
      float o[HDV];
#pragma unroll
      for (int i = 0; i < HDV; ++i) o[i] = 0.f;
      uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
      const int sw = r & 7;

      for (int j = 0; j < nblk; ++j) {
        const int s = st_of(j);
        const int kbase = kv_lo + j * ATT_BN;
        mbar_wait(&s_full[s], ph_of(j));
        tc_fence_after();
        const uint32_t tS = tmem_base + lane_off + static_cast<uint32_t>(s * ATT_BN);
        // The softmax warps are the critical path (one warp per SM sub-partition): keep the per-element instruction count
        // low and the dependency chains short -- 4 independent max / sum accumulators, scale folded into one FFMA in front of
        // MUFU.EX2, and no per-key range tests when the whole block is visible to this row (the common case away from
        // window / sample / causal boundaries).
        const bool all_visible = kbase >= lo && kbase + ATT_BN <= hi;
        // pass 1: row max (of the raw scores: scale_log2 > 0, so max commutes with the scaling)
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
        for (int c = 0; c < ATT_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + c * 32, v);
          tmem_ld_wait();
          if (all_visible) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mxa[i & 3] = fmaxf(mxa[i & 3], __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int key = kbase + c * 32 + i;
              mxa[i & 3] = fmaxf(mxa[i & 3], (key >= lo && key < hi) ? __uint_as_float(v[i]) : -INFINITY);
            }
          }
        }
        const float mx = fmaxf(m, fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * p.scale_log2);
        const float m_use = (mx == -INFINITY) ? 0.f : mx;
        const float alpha = exp2f(m - m_use);
        const float neg_m = -m_use;
        float rsa[4] = {0.f, 0.f, 0.f, 0.f};
        // pass 2: probabilities -> bf16 P tile (K-major, 128B swizzle: 16B chunk index XOR (row & 7))
#pragma unroll 1
        for (int c = 0; c < ATT_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + c * 32, v);
          tmem_ld_wait();
          float pr[32];
          if (all_visible) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
              rsa[i & 3] += pr[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int key = kbase + c * 32 + i;
              const float e = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
              pr[i] = (key >= lo && key < hi) ? e : 0.f;
              rsa[i & 3] += pr[i];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int unit = c * 4 + u;  // 16-byte unit (8 keys) within the 128-key row
            uint4 pk;
            pk.x = pack_bf16x2(pr[u * 8 + 0], pr[u * 8 + 1]);
            pk.y = pack_bf16x2(pr[u * 8 + 2], pr[u * 8 + 3]);
            pk.z = pack_bf16x2(pr[u * 8 + 4], pr[u * 8 + 5]);
            pk.w = pack_bf16x2(pr[u * 8 + 6], pr[u * 8 + 7]);
            uint8_t* dst = prow + (unit >> 3) * 16384 + (((unit & 7) ^ sw) << 4);
            *reinterpret_cast<uint4*>(dst) = pk;
          }
        }
        const float rowsum = (rsa[0] + rsa[1]) + (rsa[2] + rsa[3]);
        l = l * alpha + rowsum;
        m = mx;
        tc_fence_before();         // our TMEM reads of S / PV are ordered before the MMAs released below
        fence_proxy_async_smem();  // P (generic-proxy stores) visible to the tensor-core (async) proxy
        mbar_arrive(p_ready);
        // O = O * alpha + P V
        mbar_wait(pv_full, j & 1);
        tc_fence_after();
        const uint32_t tPV = tmem_base + lane_off + kTmemPV;
#pragma unroll
        for (int c = 0; c < HDV / 16; ++c) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tPV + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[c * 16 + i] = o[c * 16 + i] * alpha + __uint_as_float(v[i]);
        }
      }
      if (row_ok) {
        const float inv = (l > 0.f) ? 1.0f / l : 0.f;
        __nv_bfloat16* dst = p.O + static_cast<int64_t>(q_row0 + r) * p.o_row_stride + static_cast<int64_t>(head) * p.o_head_stride;
#pragma unroll
        for (int g = 0; g < HDV / 8; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv);
          pk.y = pack_bf16x2(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
          pk.z = pack_bf16x2(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv);
          pk.w = pack_bf16x2(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
          *reinterpret_cast<uint4*>(dst + g * 8) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}


static int g_attn_persistent = 1;  // SR_ATTN_PERSISTENT=0 / sr_set_attention_lean(3): one CTA per item in the variant below

// ---------------------------------------------------------------------------------------------------------------------
// "P and O in tensor memory" variant (the product configuration, SR_ATTN_LEAN = 2).  Same tiles and the same mathematics as the
// kernels above (not the same bits: the softmax reference point is lazy, see the softmax warps), different plumbing of a block:
//   * the probabilities go back into TENSOR MEMORY, over the S columns they were computed from (tcgen05.st, 2 bf16 per 32-bit
//     column), and P V is issued with the A operand in TMEM (tcgen05.mma [d], [a], b-desc).  No P tile in shared memory, no
//     generic -> async proxy fence, and -- the point -- the K tile is dead as soon as S = Q K^T has completed;
//   * K and V have their own barriers: K_{j+1} is loaded while block j is in its softmax, V is double-buffered (head_dim_v 80:
//     2 x 20 KB) and lands long before its P V.  In the lean kernel above the next block's loads started only after P V, so
//     every block paid a full L2 round trip (3.9 us per block against ~1.3 us of softmax);
//   * 104 KB (head_dim_v 80) / 96 KB (128, single V stage) of shared memory and 256 TMEM columns: still two CTAs per SM.
template <int HDV>
struct AttnTpCfg {
  static constexpr int kVStages = (HDV <= 80) ? 2 : 1;
  static constexpr int kQBytes = ATT_BM * ATT_D * 2;
  static constexpr int kKBytes = ATT_BN * ATT_D * 2;
  static constexpr int kVBoxBytes = HDV * 128;
  static constexpr int kVBytes = 2 * kVBoxBytes;
  static constexpr int kVStage = ((kVBytes + 1023) / 1024) * 1024;
  static constexpr int kSmemBytes = kQBytes + kKBytes + kVStages * kVStage + 1024 + 256;
  static constexpr int kTmemCols = 256;
  static constexpr uint32_t kTmemPV = 128;
};

SR_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
SR_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: A is [128 lanes = rows] x [K = 16 bf16 = 8 columns], K-major
SR_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int HDV>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Cfg = AttnTpCfg<HDV>;
  constexpr int VS = Cfg::kVStages;
  static_assert(HDV % 16 == 0 && HDV <= 128, "HDV");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kQBytes;
  uint8_t* sV = sK + Cfg::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + VS * Cfg::kVStage);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* v_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_ready = bars + 8;
  uint64_t* pv_full = bars + 9;
  uint64_t* q_empty = bars + 10;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Work items = (tile, head), HEAD fastest: the CTAs resident at any moment work on all heads of a few tiles, so that their Q / K
  // row reads (256 B per head out of a 4 KB token row) add up to whole rows and q heads of one kv head share its K / V tiles
  // in L2.  CTA c takes items c, c + gridDim.x, ...  The host launches either one CTA per
  // item or a PERSISTENT grid of two CTAs per SM (many short items: the single-block ViT window tiles spent most of their
  // 9 us CTA lifetime in set-up, first loads and tear-down).  All barrier phases run on counters that continue across items.
  const int n_items = p.n_tiles * p.n_heads;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(&v_full[0], 1);
    mbar_init(&v_full[1], 1);
    mbar_init(&v_empty[0], 1);
    mbar_init(&v_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(pv_full, 1);
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kTmemPV = Cfg::kTmemPV;

  {
    if (warp_idx == 0) {
      // ---------------- TMA producer ----------------
      if (lane == 0) {
        uint32_t nq = 0, nk = 0, nv = 0;  // Q tiles / K blocks / V blocks loaded so far
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
          const int4 tile = p.tiles[item / p.n_heads];
          const int head = item % p.n_heads;
          const int kv_head = head / p.group;
          const int q_row0 = tile.x, kv_lo = tile.z & ~7, kv_hi = tile.w;  // see attn_fwd_kernel
          const int nblk = (kv_hi - kv_lo + ATT_BN - 1) / ATT_BN;
          if (nblk <= 0) continue;
          mbar_wait(q_empty, (nq & 1) ^ 1);             // the previous item's last S = Q K^T has completed
          mbar_arrive_expect_tx(q_full, Cfg::kQBytes);
          tma_load_3d(sQ, &tmQ, q_full, 0, head, q_row0);
          tma_load_3d(sQ + 16384, &tmQ, q_full, 64, head, q_row0);
          ++nq;
          for (int j = 0; j < nblk; ++j) {
            const int kbase = kv_lo + j * ATT_BN;
            mbar_wait(k_empty, (nk & 1) ^ 1);           // the previous S = Q K^T has completed
            mbar_arrive_expect_tx(k_full, Cfg::kKBytes);
            tma_load_3d(sK, &tmK, k_full, 0, kv_head, kbase);
            tma_load_3d(sK + 16384, &tmK, k_full, 64, kv_head, kbase);
            ++nk;
            const int vs = nv % VS;
            mbar_wait(&v_empty[vs], ((nv / VS) & 1) ^ 1);  // P V of the block that used this V stage has completed
            mbar_arrive_expect_tx(&v_full[vs], Cfg::kVBytes);
            uint8_t* dv = sV + vs * Cfg::kVStage;
            tma_load_2d(dv, &tmV, &v_full[vs], kbase, kv_head * HDV);
            tma_load_2d(dv + Cfg::kVBoxBytes, &tmV, &v_full[vs], kbase + 64, kv_head * HDV);
            ++nv;
          }
        }
      }
    } else if (warp_idx == 1) {
      // ---------------- MMA issuer ----------------
      if (lane == 0) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, HDV);
        const uint32_t aQ = smem_u32(sQ);
        const uint32_t aK = smem_u32(sK);
        uint32_t nq = 0, nb = 0;  // items / blocks issued so far
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
          const int4 tile = p.tiles[item / p.n_heads];
          const int nblk = (tile.w - (tile.z & ~7) + ATT_BN - 1) / ATT_BN;
          if (nblk <= 0) continue;
          att_stamp(p, static_cast<int>(nq), 8);
          mbar_wait(q_full, nq & 1);
          ++nq;
          for (int j = 0; j < nblk; ++j, ++nb) {
            // S_j overwrites the columns the previous P lives in: it is issued after the previous P V (the tensor pipe runs one
            // thread's MMAs in issue order), and the softmax warps were done with the previous S before they signalled p_ready
            mbar_wait(k_full, nb & 1);
            tc_fence_after();
            if (j == 0) att_stamp(p, static_cast<int>(nq) - 1, 9);
#pragma unroll
            for (int k = 0; k < ATT_D / 16; ++k) {
              const uint64_t da = umma_desc_k_sw128(aQ + (k >> 2) * 16384) + static_cast<uint64_t>((k & 3) * 2);
              const uint64_t db = umma_desc_k_sw128(aK + (k >> 2) * 16384) + static_cast<uint64_t>((k & 3) * 2);
              umma_bf16_ss(tmem_base, da, db, idesc_s, k != 0 ? 1u : 0u);
            }
            umma_commit(s_full);
            umma_commit(k_empty);
            if (j + 1 == nblk) umma_commit(q_empty);
            const int vs = nb % VS;
            mbar_wait(&v_full[vs], (nb / VS) & 1);
            // the first P V of an item overwrites O: the softmax warps signal p_ready for it only after they have read the
            // previous item's O (program order in each of the 128 threads)
            if (j == 0) att_stamp(p, static_cast<int>(nq) - 1, 10);
            mbar_wait(p_ready, nb & 1);
            tc_fence_after();
            if (j == 0) att_stamp(p, static_cast<int>(nq) - 1, 11);
            const uint32_t aV = smem_u32(sV + vs * Cfg::kVStage);
#pragma unroll
            for (int k = 0; k < ATT_BN / 16; ++k) {
              const uint64_t db = umma_desc_k_sw128(aV + (k >> 2) * Cfg::kVBoxBytes) + static_cast<uint64_t>((k & 3) * 2);
              umma_bf16_ts(tmem_base + kTmemPV, tmem_base + static_cast<uint32_t>(k * 8), db, idesc_pv, (j | k) != 0 ? 1u : 0u);
            }
            umma_commit(&v_empty[vs]);
            if (j + 1 == nblk) umma_commit(pv_full);
          }
        }
      }
    } else {
      // ---------------- softmax / output warps: one query row per thread ----------------
      // O is accumulated by the tensor core IN tensor memory (P V with accumulate); this thread keeps only (m, l) of its row
      // and the whole 128-key score row in registers: one TMEM round trip per block instead of 13 (two passes of four chunk
      // loads + five O chunk loads in the kernels above).  The running maximum is LAZY: the reference point m moves only when
      // the block maximum exceeds it by more than 2^8 (probabilities then stay <= 256, harmless in fp32 / bf16), in which case
      // the warp rescales its O rows in TMEM -- rare after the first block, so O normally never leaves tensor memory until
      // the epilogue.  S_j complete implies P V_{j-1} complete (one thread's MMAs finish in issue order), so O may be touched
      // as soon as s_full fires.
      const int quarter = warp_idx & 3;
      const int r = quarter * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t tS = tmem_base + lane_off;
      const uint32_t tPV = tmem_base + lane_off + kTmemPV;
      uint32_t nb = 0, ni = 0;  // blocks / items processed so far
      // this row's key range of the NEXT item is fetched one item ahead (an L2 round trip at the head of every item otherwise)
      auto row_range = [&](int item, int& lo_, int& hi_) {
        lo_ = hi_ = 0;
        if (item < n_items) {
          const int4 t = p.tiles[item / p.n_heads];
          if (r < t.y) {
            lo_ = __ldg(p.row_lo + t.x + r);
            hi_ = __ldg(p.row_hi + t.x + r);
          }
        }
      };
      int lo_next, hi_next;
      row_range(blockIdx.x, lo_next, hi_next);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int4 tile = p.tiles[item / p.n_heads];
      const int head = item % p.n_heads;
      const int q_row0 = tile.x, q_nrows = tile.y, kv_lo = tile.z & ~7, kv_hi = tile.w;
      const int nblk = (kv_hi - kv_lo + ATT_BN - 1) / ATT_BN;
      const bool row_ok = r < q_nrows;
      const int lo = lo_next, hi = hi_next;
      row_range(item + gridDim.x, lo_next, hi_next);
      if (nblk <= 0) continue;
      float m = -INFINITY, l = 0.f;
      const bool stamp = r == 0;
      if (stamp) att_stamp(p, static_cast<int>(ni), 0);

      for (int j = 0; j < nblk; ++j, ++nb) {
        const int kbase = kv_lo + j * ATT_BN;
        mbar_wait(s_full, nb & 1);
        tc_fence_after();
        if (stamp && j == 0) att_stamp(p, static_cast<int>(ni), 1);
        uint32_t v0[32], v1[32], v2[32], v3[32];  // the row's 128 scores (separate arrays: they must stay in registers)
        tmem_ld_32x32b_x32(tS, v0);
        tmem_ld_32x32b_x32(tS + 32, v1);
        tmem_ld_32x32b_x32(tS + 64, v2);
        tmem_ld_32x32b_x32(tS + 96, v3);
        tmem_ld_wait();
        if (stamp && j == 0) att_stamp(p, static_cast<int>(ni), 2);
        const bool all_visible = kbase >= lo && kbase + ATT_BN <= hi;
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto max_chunk = [&](uint32_t(&vv)[32], int c) {
          if (!all_visible) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int key = kbase + c * 32 + i;
              if (!(key >= lo && key < hi)) vv[i] = 0xff800000u;  // -inf: exp2 -> 0, never the maximum
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) mxa[i & 3] = fmaxf(mxa[i & 3], __uint_as_float(vv[i]));
        };
        max_chunk(v0, 0);
        max_chunk(v1, 1);
        max_chunk(v2, 2);
        max_chunk(v3, 3);
        const float mblk = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * p.scale_log2;
        float alpha = 1.f;
        bool moved = false;
        if (m == -INFINITY) {
          m = mblk;                       // nothing accumulated for this row yet (O row and l are 0)
        } else if (mblk > m + 8.f) {
          alpha = exp2f(m - mblk);
          m = mblk;
          moved = true;
        }
        if (j > 0 && __any_sync(0xffffffffu, moved)) {
#pragma unroll
          for (int c = 0; c < HDV / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(tPV + c * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32b_x16(tPV + c * 16, ov);
          }
        }
        if (stamp && j == 0) att_stamp(p, static_cast<int>(ni), 3);
        const float neg_m = (m == -INFINITY) ? 0.f : -m;
        float rsa[4] = {0.f, 0.f, 0.f, 0.f};
        auto exp_chunk = [&](const uint32_t(&vv)[32], int c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(vv[2 * i]), p.scale_log2, neg_m));
            const float p1 = ex2_approx(fmaf(__uint_as_float(vv[2 * i + 1]), p.scale_log2, neg_m));
            rsa[i & 3] += p0 + p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x16(tS + c * 16, pk);
        };
        exp_chunk(v0, 0);
        exp_chunk(v1, 1);
        exp_chunk(v2, 2);
        exp_chunk(v3, 3);
        l = l * alpha + ((rsa[0] + rsa[1]) + (rsa[2] + rsa[3]));
        if (stamp && j == 0) att_stamp(p, static_cast<int>(ni), 4);
        tmem_st_wait();
        tc_fence_before();  // our TMEM accesses (S reads, P / O stores) are ordered before the MMAs released below
        mbar_arrive(p_ready);
        if (stamp && j == 0) att_stamp(p, static_cast<int>(ni), 5);
      }
      mbar_wait(pv_full, ni & 1);  // committed once per item, after its last P V
      if (stamp) att_stamp(p, static_cast<int>(ni), 6);
      ++ni;
      tc_fence_after();
      {
        const float inv = (l > 0.f) ? 1.0f / l : 0.f;
        __nv_bfloat16* dst = p.O + static_cast<int64_t>(q_row0 + r) * p.o_row_stride + static_cast<int64_t>(head) * p.o_head_stride;
#pragma unroll
        for (int c = 0; c < HDV / 16; ++c) {
          uint32_t ov[16];
          tmem_ld_32x32b_x16(tPV + c * 16, ov);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
              pk.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
              pk.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
              pk.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + c * 16 + g * 8) = pk;
            }
          }
        }
      }
      if (stamp) att_stamp(p, static_cast<int>(ni) - 1, 7);
      tc_fence_before();  // the O reads above are ordered before the next item's first P V (released through p_ready)
      }  // items
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int HDV>
static int launch_attn_tp(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnParams& p, int n_tiles,
                          int n_heads, cudaStream_t stream) {
  using Cfg = AttnTpCfg<HDV>;
  auto kern = attn_fwd_tp_kernel<HDV>;
  static unsigned long long attr_done = 0;  // per device
  SR_CHECK_CUDA(ensure_dynamic_smem(kern, Cfg::kSmemBytes, &attr_done));
  // Persistent grid (two CTAs per SM, static round-robin over the items) when every CTA gets many items -- the ViT plans, whose
  // items are equally long; the causal prefill plan (a few items of 1 .. 7 blocks per CTA) keeps one CTA per item and leaves
  // the balancing to the hardware scheduler (tiles are sorted longest-first).
  const int n_items = n_tiles * n_heads;
  const int resident = 2 * sm_count();
  const int grid = (g_attn_persistent && n_items >= 6 * resident) ? resident : n_items;
  kern<<<grid, ATT_THREADS, Cfg::kSmemBytes, stream>>>(q, k, v, p);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

template <int HDV, int STAGES>
static int launch_attn(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnParams& p, int n_tiles,
                       int n_heads, cudaStream_t stream) {
  using Cfg = AttnCfg<HDV, STAGES>;
  auto kern = attn_fwd_kernel<HDV, STAGES>;
  static unsigned long long attr_done = 0;  // per device
  SR_CHECK_CUDA(ensure_dynamic_smem(kern, Cfg::kSmemBytes, &attr_done));
  dim3 grid(n_tiles, n_heads);
  kern<<<grid, ATT_THREADS, Cfg::kSmemBytes, stream>>>(q, k, v, p);
  SR_CHECK_CUDA(cudaGetLastError());
  return SR_OK;
}

static long long* g_attn_trace = nullptr;  // sr_set_attention_trace (diagnostics)
static int g_attn_lean = -1;  // -1: env SR_ATTN_LEAN (default 2); 0 = one CTA / SM, 1 = lean, 2 = lean with P in tensor memory

}  // namespace sr

extern "C" int sr_set_attention_trace(void* buffer) {
  sr::g_attn_trace = reinterpret_cast<long long*>(buffer);
  return SR_OK;
}

extern "C" int sr_set_attention_lean(int enabled) {
  // 0 / 1 / 2: kernel variant (see g_attn_lean); 3: variant 2 with one CTA per work item instead of the persistent grid
  sr::g_attn_persistent = enabled == 3 ? 0 : 1;
  sr::g_attn_lean = enabled < 0 ? -1 : (enabled > 2 ? 2 : enabled);
  return SR_OK;
}

extern "C" int sr_attention_fwd(const void* Q, int64_t q_row_stride, int64_t q_head_stride, int n_q_rows, int n_q_heads,
                                const void* K, int64_t k_row_stride, int64_t k_head_stride, int n_k_rows, int n_kv_heads,
                                const void* Vt, int64_t vt_row_stride, int n_vt_cols, void* O, int64_t o_row_stride,
                                int64_t o_head_stride, const int32_t* tiles, int n_tiles, const int32_t* row_lo,
                                const int32_t* row_hi, float scale, int head_dim_v, void* stream) {
  using namespace sr;
  SR_REQUIRE(Q && K && Vt && O && tiles && row_lo && row_hi, "sr_attention_fwd: null pointer");
  SR_REQUIRE(n_tiles > 0 && n_q_heads > 0 && n_kv_heads > 0 && n_q_heads % n_kv_heads == 0, "sr_attention_fwd: bad head/tile counts");
  SR_REQUIRE(head_dim_v == 80 || head_dim_v == 128, "sr_attention_fwd: head_dim_v must be 80 or 128");
  SR_REQUIRE(q_row_stride % 8 == 0 && q_head_stride % 8 == 0 && k_row_stride % 8 == 0 && k_head_stride % 8 == 0 &&
                 vt_row_stride % 8 == 0 && o_row_stride % 8 == 0 && o_head_stride % 8 == 0,
             "sr_attention_fwd: strides must be multiples of 8 elements");
  SR_REQUIRE((reinterpret_cast<uintptr_t>(O) & 15) == 0 && (reinterpret_cast<uintptr_t>(tiles) & 15) == 0,
             "sr_attention_fwd: O / tiles must be 16B aligned");
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)ATT_D, (uint64_t)n_q_heads, (uint64_t)n_q_rows};
    uint64_t strides[2] = {(uint64_t)q_head_stride * 2, (uint64_t)q_row_stride * 2};
    uint32_t box[3] = {64, 1, (uint32_t)ATT_BM};
    int rc = encode_tmap_bf16(&tmQ, Q, 3, dims, strides, box, 128);
    if (rc != SR_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)ATT_D, (uint64_t)n_kv_heads, (uint64_t)n_k_rows};
    uint64_t strides[2] = {(uint64_t)k_head_stride * 2, (uint64_t)k_row_stride * 2};
    uint32_t box[3] = {64, 1, (uint32_t)ATT_BN};
    int rc = encode_tmap_bf16(&tmK, K, 3, dims, strides, box, 128);
    if (rc != SR_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)n_vt_cols, (uint64_t)n_kv_heads * (uint64_t)head_dim_v};
    uint64_t strides[1] = {(uint64_t)vt_row_stride * 2};
    uint32_t box[2] = {64, (uint32_t)head_dim_v};
    int rc = encode_tmap_bf16(&tmV, Vt, 2, dims, strides, box, 128);
    if (rc != SR_OK) return rc;
  }
  AttnParams p;
  p.O = reinterpret_cast<__nv_bfloat16*>(O);
  p.o_row_stride = o_row_stride;
  p.o_head_stride = o_head_stride;
  p.tiles = reinterpret_cast<const int4*>(tiles);
  p.row_lo = row_lo;
  p.row_hi = row_hi;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.group = n_q_heads / n_kv_heads;
  p.n_tiles = n_tiles;
  p.n_heads = n_q_heads;
  p.trace = g_attn_trace;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (g_attn_lean < 0) {
    const char* e = getenv("SR_ATTN_LEAN");
    g_attn_lean = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? (e[0] - '0') : 2;
  }
  if (g_attn_lean == 2) {
    if (head_dim_v == 80) return launch_attn_tp<80>(tmQ, tmK, tmV, p, n_tiles, n_q_heads, s);
    return launch_attn_tp<128>(tmQ, tmK, tmV, p, n_tiles, n_q_heads, s);
  }
  if (g_attn_lean) {
    if (head_dim_v == 80) return launch_attn<80, 1>(tmQ, tmK, tmV, p, n_tiles, n_q_heads, s);
    return launch_attn<128, 1>(tmQ, tmK, tmV, p, n_tiles, n_q_heads, s);
  }
  if (head_dim_v == 80) return launch_attn<80, 2>(tmQ, tmK, tmV, p, n_tiles, n_q_heads, s);
  return launch_attn<128, 2>(tmQ, tmK, tmV, p, n_tiles, n_q_heads, s);
}
