// sr_decode_persistent.cu -- one decode step (all decoder layers + final norm + LM head) as ONE persistent kernel.
//
// What it replaces: the per-token body of the reference's rollout decode loop (vLLM engine step,
// roll/distributed/strategy/vllm_strategy.py:114-141; HF eager equivalent: Qwen2_5_VLDecoderLayer x L with use_cache,
// HF:762-827, 672-759, 611-624, final norm + lm_head HF:1385) -- here for M <= 8 live sequences.
//
// Why: at M = 8 the step is pure weight streaming (6.2 GB for the 3B model).  Launched as ~6 kernels per layer, every kernel
// pays a ramp (first bytes arrive after one HBM latency) and a tail (cross-warp reduction + epilogue with no loads in flight)
// -- measured 2.84 ms/token against ~1.0 ms of streaming at the measured HBM rate.  Here the HBM stream
// never stops:
//   * grid = one CTA per SM, resident for the whole step (cooperative launch);
//   * warp 0 of every CTA is a PRODUCER: one thread walks the static schedule (layer -> {qkv, o, gate/up, down} -> this CTA's
//     row groups -> 16 KB stages) and issues cp.async.bulk copies of the pre-packed weights ([N/16][K/32][16][32], 1 KB
//     blocks) into a 10-stage / 160 KB shared-memory ring guarded by full/empty mbarriers.  Weights never depend on
//     activations, so the producer runs AHEAD of the consumers across phase boundaries and grid barriers: while the CTAs
//     synchronise, do attention or run an epilogue, ~24 MB of the next matrices are already in flight chip-wide;
//   * warps 1-8 are CONSUMERS: per phase they stage the 8 activation rows in shared memory (RMSNorm fused in the staging,
//     HF:57-71), read 16-byte weight fragments from the ring straight into mma.sync.m16n8k16 A operands (same k permutation
//     as sr_gemv.cu: a 16 B load IS a fragment), reduce across warps, run the fused epilogue (bias / residual / SwiGLU);
//   * phases are separated by a grid barrier (release-add + acquire-poll on one counter): qkv | rope + KV append + split-KV
//     attention (flash-decoding, merge by the last CTA of a (sample, kv head)) | o-proj + residual | gate/up + SwiGLU |
//     down + residual; then final norm + LM head (fp32 logits).
// HBM-bound: algorithmic bytes per launch = all decoder weights + LM head (N*K*2 each) + KV cache of the live sequences.
#include <stdio.h>
#include <stdlib.h>

#include "sr_host.h"

#ifndef SR_DP_NOINLINE_BAR
#define SR_DP_NOINLINE_BAR 0  // measured: 2.07 ms/step inlined vs 2.16 as a call
#endif
#ifndef SR_DP_NOINLINE_STAGE
#define SR_DP_NOINLINE_STAGE 0
#endif
#if SR_DP_NOINLINE_BAR
#define DP_BAR_ATTR __device__ __noinline__
#else
#define DP_BAR_ATTR SR_DEVINL
#endif
#if SR_DP_NOINLINE_STAGE
#define DP_STAGE_ATTR __device__ __noinline__
#else
#define DP_STAGE_ATTR SR_DEVINL
#endif

namespace sr {

constexpr int DP_CONSUMERS = 8;
constexpr int DP_CTHREADS = DP_CONSUMERS * 32;
constexpr int DP_THREADS = DP_CTHREADS + 64;          // warp 0 = producer, warps 1-8 = consumers, warp 9 = L2 prefetcher
constexpr int DP_STAGE_BLOCKS = 16;                   // 1 KB weight blocks per ring stage
constexpr int DP_STAGE_BYTES = DP_STAGE_BLOCKS * 1024;
constexpr int DP_STAGES = 10;
constexpr int DP_H = 2048;                            // staged activation width (hidden size = n_q_heads * 128)
constexpr int DP_XS_STRIDE = DP_H * 2 + 64;           // bytes per staged row (+64 B: conflict-free 16 B fragment reads)
constexpr int DP_XPF = 4;                             // activation prefetch distance (stages) on the L2 path (loaded L2 latency ~1.5 us)
constexpr int DP_MAX_LAYERS = 64;                     // layer table staged in shared memory
constexpr int DP_GMAX = 8;                            // query heads per kv head (max)
constexpr int DP_KEYS = 128;                          // keys per attention pass
constexpr int DP_NS_MAX = 32;                         // max key splits
constexpr int DP_PART = DP_GMAX * 128 + 2 * DP_GMAX;  // floats per attention partial (o, m, l)

// shared-memory carve-up (bytes, from a 1024-aligned base)
constexpr int DP_OFF_RING = 0;
constexpr int DP_OFF_XS = DP_STAGES * DP_STAGE_BYTES;                 // staged activations; aliased by attention scratch
constexpr int DP_XS_BYTES = 8 * DP_XS_STRIDE;                         // 33280
constexpr int DP_OFF_RED = DP_OFF_XS + DP_XS_BYTES;                   // [2][8 warps][2 rg][128] fp32
constexpr int DP_RED_BYTES = 2 * DP_CONSUMERS * 2 * 128 * 4;          // 16384
constexpr int DP_OFF_BAR = DP_OFF_RED + DP_RED_BYTES;
constexpr int DP_OFF_LAYERS = DP_OFF_BAR + 2 * DP_STAGES * 8 + 32;    // DpLayer[DP_MAX_LAYERS]
constexpr int DP_OFF_ROPE = DP_OFF_LAYERS + DP_MAX_LAYERS * 56;       // cos[8][64], sin[8][64] fp32 of this step
constexpr int DP_OFF_SCAL = DP_OFF_ROPE + 2 * 8 * 64 * 4;             // kv_start[8], kv_len[8], cache_row[8]
constexpr int DP_OFF_PCNT = DP_OFF_SCAL + 3 * 8 * 4;                  // stages issued by the producer so far (read by the prefetcher)
constexpr int DP_SMEM_BYTES = DP_OFF_PCNT + 64 + 1024;                // + alignment slack
static_assert(DP_SMEM_BYTES <= 232448, "shared-memory budget of one sm_100 CTA");


struct DpLayer {  // == sr_decode_layer (include/socioreasoner_b200.h)
  const __nv_bfloat16* qkv_wp;
  const float* qkv_b;
  const float* ln1;
  const __nv_bfloat16* o_wp;
  const float* ln2;
  const __nv_bfloat16* gu_wp;
  const __nv_bfloat16* down_wp;
};

struct DpParams {
  const DpLayer* layers;
  int n_layers;
  const __nv_bfloat16* lm_head_p;
  const float* final_norm;
  int vocab;
  __nv_bfloat16* x;
  int64_t ldx;
  __nv_bfloat16* qkv;      // [B, (nq + 2 nkv) * 128]
  __nv_bfloat16* ao;       // [B, nq * 128]
  __nv_bfloat16* act;      // [B, I]
  float* logits;
  int64_t ld_logits;
  const float* cosb;       // [B, 64]
  const float* sinb;
  const int* cache_row;
  const int* kv_start;
  const int* kv_len;
  __nv_bfloat16* Kc;       // [n_layers][rows, nkv, 128]
  int64_t k_layer_stride;
  __nv_bfloat16* Vt;       // [n_layers][nkv*128, ldvt]
  int64_t vt_layer_stride;
  int64_t ldvt;
  int B, nq, nkv, I;
  float eps, scale_log2;
  float* attn_ws;
  int* attn_tickets;
  int ns;
  unsigned int* grid_bar;  // zero before launch
  int stop_after;          // debug: run only the first `stop_after` phases (<= 0: all)
  int pf_stages;           // L2 prefetch lookahead of the producer beyond the ring, in 16 KB stages per CTA (0 = off)
  int ring_stages;         // shared-memory ring stages in use (2 .. DP_STAGES)
  int unit_off[4];         // unit -> CTA rotation of the qkv / o / gate-up / down matrices
  int flags;               // bit 0: L2 evict_first hint on the ring copies (weights are consumed exactly once)
  unsigned long long* trace;  // optional [gridDim.x][2 roles][trace_events][2] (globaltimer ns, clock64) -- sr_set_decode_persistent_trace
  int trace_events;
};

// In-kernel timeline (bring-up / tuning aid): thread-level timestamps at phase boundaries, read back by tools/persistent_trace.py.
SR_DEVINL void dp_trace_at(const DpParams& p, int role, int idx) {
  if (p.trace != nullptr && idx < p.trace_events) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    unsigned long long* d = p.trace + ((static_cast<int64_t>(blockIdx.x) * 2 + role) * p.trace_events + idx) * 2;
    d[0] = t;
    d[1] = static_cast<unsigned long long>(clock64());
  }
}
SR_DEVINL void dp_trace(const DpParams& p, int role, int& idx) {
  if (p.trace != nullptr && idx < p.trace_events) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    unsigned long long* d = p.trace + ((static_cast<int64_t>(blockIdx.x) * 2 + role) * p.trace_events + idx) * 2;
    d[0] = t;
    d[1] = static_cast<unsigned long long>(clock64());
    ++idx;
  }
}

SR_DEVINL void dp_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
SR_DEVINL void dp_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
SR_DEVINL void dp_cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // consumer warps only
SR_DEVINL uint4 dp_ldcg(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

#ifndef SR_GRIDBAR_SPIN_LIMIT
#define SR_GRIDBAR_SPIN_LIMIT (1u << 23)
#endif
// Grid barrier for the consumer threads of all CTAs.  Arrive: the CTA barrier orders every consumer thread's global stores
// before thread 32's red.release (cumulativity).  Wait: relaxed polls, then ONE ld.acquire of the counter -- it reads a value of
// the RMW chain that contains every CTA's release-increment of this generation, so it synchronises with all of them; the second
// CTA barrier extends that to the other threads.  (Measured: an acq_rel fence after the poll costs 0.3 us per barrier more than
// the acquire load, 2.02 vs 1.96 ms per step; dropping the release too would save another 0.4 us but is not memory-model safe.)
// Bounded: a broken schedule traps (CUDA error through the C ABI) instead of hanging the GPU.
DP_BAR_ATTR void dp_grid_barrier(unsigned int* ctr, unsigned int target) {
  dp_cbar();
  if (threadIdx.x == 32) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    unsigned int v = 0, spins = 0;
    while (true) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (++spins > SR_GRIDBAR_SPIN_LIMIT) {
        printf("sr: grid barrier timeout block=%d have=%u want=%u\n", blockIdx.x, v, target);
        __trap();
      }
    }
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
  }
  dp_cbar();
}

struct DpRing {
  uint8_t* data;
  uint64_t* full;
  uint64_t* empty;
  int slot;
  uint32_t phase;
  int n;  // stages in use (<= DP_STAGES): fewer stages = less lookahead but also a shorter queue in front of the latency chains
  SR_DEVINL void advance() {
    if (++slot == n) { slot = 0; phase ^= 1; }
  }
};

// ------------------------------------------------------------------------------------------------ producer
// Streams this CTA's units of one packed matrix: unit u = RG consecutive 16-row groups; stage = 16 blocks of 1 KB
// (RG = 1: 16 k-chunks of the group; RG = 2: 8 k-chunks of each of the two groups, second group at +8 KB).
// ONE thread issues every copy of the CTA, so this loop is kept minimal (a generic schedule walker with an L2-prefetch
// cursor was measured 15% slower end to end: the producer thread became the bottleneck).
// ROT: the stage order of unit u is rotated by (7u mod nst) -- used for the down projection, whose activations are read
// from L2 by every CTA: without the rotation all CTAs request the same 8 KB of activations at the same time (hot L2 slices).
SR_DEVINL int dp_rot(int u, int nst) { return (u * 7) % nst; }
template <int RG, bool ROT>
SR_DEVINL void dp_produce(DpRing& r, int v0, const __nv_bfloat16* Wp, int n_units, int nchunks, uint64_t policy, bool hint,
                          volatile int* pcnt, int& issued) {
  constexpr int CPS = DP_STAGE_BLOCKS / RG;
  const int nst = (nchunks + CPS - 1) / CPS;
  for (int u = v0; u < n_units; u += gridDim.x) {
    int st = ROT ? dp_rot(u, nst) : 0;
    for (int i = 0; i < nst; ++i) {
      mbar_wait(&r.empty[r.slot], r.phase ^ 1);
      const int nch = min(CPS, nchunks - st * CPS);
      mbar_arrive_expect_tx(&r.full[r.slot], static_cast<uint32_t>(RG * nch * 1024));
#pragma unroll
      for (int g = 0; g < RG; ++g) {
        const __nv_bfloat16* src = Wp + (static_cast<int64_t>(u * RG + g) * nchunks + st * CPS) * 512;
        uint8_t* dst = r.data + r.slot * DP_STAGE_BYTES + g * (CPS * 1024);
        if (hint) {
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                  smem_u32(dst)),
              "l"(src), "r"(static_cast<uint32_t>(nch * 1024)), "r"(smem_u32(&r.full[r.slot])), "l"(policy)
              : "memory");
        } else {
          dp_bulk_g2s(dst, src, static_cast<uint32_t>(nch * 1024), &r.full[r.slot]);
        }
      }
      r.advance();
      *pcnt = ++issued;
      if (ROT && ++st == nst) st = 0;
      if (!ROT) ++st;
    }
  }
}

// L2 prefetcher (warp 9, one thread): walks the SAME stage sequence as the producer, at most `lead` stages ahead of it, and asks
// the memory system to bring each stage into L2 (cp.async.bulk.prefetch.L2).  The shared-memory ring holds only 160 KB per CTA
// (3.7 us of HBM time chip-wide), and the consumers spend most of a layer in latency chains (qkv -> attention -> o-proj: grid
// barriers, L2 round trips) during which the ring is full and HBM would idle; with the prefetcher the HBM stream continues into
// L2 (`lead` x 16 KB x 148 CTAs), and the ring later refills at L2 speed.
template <int RG, bool ROT>
SR_DEVINL void dp_prefetch_l2(int& sidx, const volatile int* pcnt, int lead, int v0, const __nv_bfloat16* Wp, int n_units, int nchunks,
                               uint64_t policy, bool hint) {
  constexpr int CPS = DP_STAGE_BLOCKS / RG;
  const int nst = (nchunks + CPS - 1) / CPS;
  for (int u = v0; u < n_units; u += gridDim.x) {
    int st = ROT ? dp_rot(u, nst) : 0;
    for (int i = 0; i < nst; ++i) {
      int pc = *pcnt;
      while (sidx >= pc + lead) {
        __nanosleep(100);
        pc = *pcnt;
      }
      if (sidx > pc) {  // (the producer is about to request stage `pc` itself)
        const int nch = min(CPS, nchunks - st * CPS);
#pragma unroll
        for (int g = 0; g < RG; ++g) {
          const __nv_bfloat16* src = Wp + (static_cast<int64_t>(u * RG + g) * nchunks + st * CPS) * 512;
          if (hint)
            asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(static_cast<uint32_t>(nch * 1024)),
                         "l"(policy)
                         : "memory");
          else
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(static_cast<uint32_t>(nch * 1024)) : "memory");
        }
      }
      ++sidx;
      if (ROT && ++st == nst) st = 0;
      if (!ROT) ++st;
    }
  }
}

// ------------------------------------------------------------------------------------------------ consumers
// Stage M (<= 8) rows of width DP_H into shared memory, warp cw <-> row cw, optional RMSNorm (y = w * x * rsqrt(mean x^2 + eps),
// one rounding to bf16 -- the arithmetic of rmsnorm_kernel in sr_elementwise.cu).  Rows >= M are zero.
template <bool NORM>
DP_STAGE_ATTR void dp_stage_rows(uint8_t* xs, const __nv_bfloat16* src, int64_t ld, int M, const float* nw, float eps, int cw,
                             int lane) {
  constexpr int NV = DP_H / 256;  // 16 B vectors per lane
  uint4 v[NV];
  float4 nw0[NORM ? NV : 1], nw1[NORM ? NV : 1];
  const bool ok = cw < M;
  if (NORM) {  // the norm weights do not depend on x: requested together with it (one round trip instead of two)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      nw0[i] = __ldg(reinterpret_cast<const float4*>(nw + (lane + 32 * i) * 8));
      nw1[i] = __ldg(reinterpret_cast<const float4*>(nw + (lane + 32 * i) * 8 + 4));
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
    v[i] = ok ? dp_ldcg(src + static_cast<int64_t>(cw) * ld + (lane + 32 * i) * 8) : make_uint4(0, 0, 0, 0);
  if (NORM) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
        ss += a * a + b * b;
      }
    }
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss / static_cast<float>(DP_H) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 w0 = nw0[i], w1 = nw1[i];
      uint4 o;
      o.x = pack_bf16x2(bf16_lo(v[i].x) * inv * w0.x, bf16_hi(v[i].x) * inv * w0.y);
      o.y = pack_bf16x2(bf16_lo(v[i].y) * inv * w0.z, bf16_hi(v[i].y) * inv * w0.w);
      o.z = pack_bf16x2(bf16_lo(v[i].z) * inv * w1.x, bf16_hi(v[i].z) * inv * w1.y);
      o.w = pack_bf16x2(bf16_lo(v[i].w) * inv * w1.z, bf16_hi(v[i].w) * inv * w1.w);
      v[i] = o;
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) *reinterpret_cast<uint4*>(xs + cw * DP_XS_STRIDE + (lane + 32 * i) * 16) = v[i];
}

// K loop of one unit with the activations staged in shared memory (K = DP_H).
template <int RG>
SR_DEVINL void dp_consume_staged(DpRing& r, int nchunks, const uint8_t* xs, float (&acc)[RG][4], int cw, int lane) {
  constexpr int CPS = DP_STAGE_BLOCKS / RG;
  const int g = lane >> 2, t = lane & 3;
  const int nst = (nchunks + CPS - 1) / CPS;
  const uint8_t* xrow = xs + g * DP_XS_STRIDE + t * 16;
  for (int st = 0; st < nst; ++st) {
    mbar_wait(&r.full[r.slot], r.phase);
    const uint8_t* stage = r.data + r.slot * DP_STAGE_BYTES;
    if (RG == 2) {
      const int c = st * CPS + cw;
      if (c < nchunks) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xrow + c * 64);
#pragma unroll
        for (int q = 0; q < RG; ++q) {
          const uint8_t* blk = stage + q * (CPS * 1024) + cw * 1024;
          const uint4 w0 = *reinterpret_cast<const uint4*>(blk + g * 64 + t * 16);
          const uint4 w1 = *reinterpret_cast<const uint4*>(blk + (g + 8) * 64 + t * 16);
          dp_mma(acc[q], w0.x, w1.x, w0.y, w1.y, xv.x, xv.y);
          dp_mma(acc[q], w0.z, w1.z, w0.w, w1.w, xv.z, xv.w);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = st * CPS + cw + 8 * j;
        if (c < nchunks) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xrow + c * 64);
          const uint8_t* blk = stage + (cw + 8 * j) * 1024;
          const uint4 w0 = *reinterpret_cast<const uint4*>(blk + g * 64 + t * 16);
          const uint4 w1 = *reinterpret_cast<const uint4*>(blk + (g + 8) * 64 + t * 16);
          dp_mma(acc[0], w0.x, w1.x, w0.y, w1.y, xv.x, xv.y);
          dp_mma(acc[0], w0.z, w1.z, w0.w, w1.w, xv.z, xv.w);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.empty[r.slot]);
    r.advance();
  }
}

// K loop of one unit (RG = 1) with the activations read from global memory / L2 (long K: the down projection).
// Stage order rotated per unit (see dp_produce<.., ROT>); activation fragments are prefetched DP_XPF stages ahead.
SR_DEVINL void dp_consume_l2(DpRing& r, int u, int nchunks, const __nv_bfloat16* xg, int64_t ldx, int M, float (&acc)[1][4], int cw,
                             int lane) {
  constexpr int CPS = DP_STAGE_BLOCKS;
  const int g = lane >> 2, t = lane & 3;
  const int nst = (nchunks + CPS - 1) / CPS;
  const int rot = dp_rot(u, nst);
  const bool xok = g < M;
  const __nv_bfloat16* xrow = xg + static_cast<int64_t>(xok ? g : 0) * ldx + t * 8;
  uint4 xq[DP_XPF][2];
  auto stage_of = [&](int i) { const int s2 = i + rot; return s2 >= nst ? s2 - nst : s2; };
  auto load_x = [&](int slot_i, int i) {
    const int st = stage_of(i);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = st * CPS + cw + 8 * j;
      xq[slot_i][j] = (xok && i < nst && c < nchunks) ? dp_ldcg(xrow + c * 32) : make_uint4(0, 0, 0, 0);
    }
  };
#pragma unroll
  for (int i = 0; i < DP_XPF; ++i) load_x(i, i);
  for (int i0 = 0; i0 < nst; i0 += DP_XPF) {
#pragma unroll
    for (int q = 0; q < DP_XPF; ++q) {
      const int i = i0 + q;
      if (i < nst) {
        const int st = stage_of(i);
        mbar_wait(&r.full[r.slot], r.phase);
        const uint8_t* stage = r.data + r.slot * DP_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = st * CPS + cw + 8 * j;
          if (c < nchunks) {
            const uint4 xv = xq[q][j];
            const uint8_t* blk = stage + (cw + 8 * j) * 1024;
            const uint4 w0 = *reinterpret_cast<const uint4*>(blk + g * 64 + t * 16);
            const uint4 w1 = *reinterpret_cast<const uint4*>(blk + (g + 8) * 64 + t * 16);
            dp_mma(acc[0], w0.x, w1.x, w0.y, w1.y, xv.x, xv.y);
            dp_mma(acc[0], w0.z, w1.z, w0.w, w1.w, xv.z, xv.w);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&r.empty[r.slot]);
        load_x(q, i + DP_XPF);
        r.advance();
      }
    }
  }
}

enum { DP_EPI_BIAS = 0, DP_EPI_RESIDUAL = 1, DP_EPI_SWIGLU = 2, DP_EPI_F32 = 3 };

// One GEMV phase: all units of this CTA.  `staged`: activations come from xs (K = DP_H), else from xg (L2).
template <int RG, int EPI, bool STAGED>
SR_DEVINL void dp_gemv_phase(DpRing& r, int v0, int n_units, int nchunks, const uint8_t* xs, const __nv_bfloat16* xg, int64_t ldxg,
                             int M, float* red, int& red_par, const float* bias, __nv_bfloat16* out_bf, float* out_f32,
                             int64_t ldo, int cw, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int ct = threadIdx.x - 32;
  for (int u = v0; u < n_units; u += gridDim.x) {
    float acc[RG][4];
#pragma unroll
    for (int q = 0; q < RG; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
    // epilogue operands of this thread's output element(s): requested before the K loop so their latency is hidden.
    // Units of the non-SwiGLU epilogues with RG = 2 are two independent 16-column groups (columns (u*RG + q)*16 ..).
    constexpr int NG = (EPI == DP_EPI_SWIGLU) ? 1 : RG;
    float epi_add[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
      epi_add[q] = 0.f;
      if (ct < 128 && (ct >> 4) < M) {
        const int col = (u * NG + q) * 16 + (ct & 15);
        if (EPI == DP_EPI_BIAS && bias != nullptr) epi_add[q] = __ldg(bias + col);
        if (EPI == DP_EPI_RESIDUAL) {
          const unsigned short rbits = __ldcg(reinterpret_cast<const unsigned short*>(out_bf + static_cast<int64_t>(ct >> 4) * ldo + col));
          epi_add[q] = __uint_as_float(static_cast<uint32_t>(rbits) << 16);
        }
      }
    }
    if constexpr (STAGED) {
      dp_consume_staged<RG>(r, nchunks, xs, acc, cw, lane);
    } else {
      static_assert(STAGED || RG == 1, "the L2 activation path handles single row groups");
      dp_consume_l2(r, u, nchunks, xg, ldxg, M, acc, cw, lane);
    }
    // cross-warp reduction: C fragment (feature g | g+8, batch 2t | 2t+1) -> red[par][warp][q][feature*8 + batch]
    float* rp = red + red_par * (DP_CONSUMERS * 2 * 128);
#pragma unroll
    for (int q = 0; q < RG; ++q) {
      float* d = rp + (cw * 2 + q) * 128;
      d[g * 8 + 2 * t] = acc[q][0];
      d[g * 8 + 2 * t + 1] = acc[q][1];
      d[(g + 8) * 8 + 2 * t] = acc[q][2];
      d[(g + 8) * 8 + 2 * t + 1] = acc[q][3];
    }
    dp_cbar();
    if (ct < 128) {
      const int b = ct >> 4, f = ct & 15;  // consecutive threads -> consecutive features
      float v[RG];
#pragma unroll
      for (int q = 0; q < RG; ++q) {
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < DP_CONSUMERS; ++w8) s += rp[(w8 * 2 + q) * 128 + f * 8 + b];
        v[q] = s;
      }
      if (b < M) {
        if (EPI == DP_EPI_SWIGLU) {
          const float gate = v[0], up = v[RG - 1];
          const float o = (gate / (1.0f + __expf(-gate))) * up;
          out_bf[static_cast<int64_t>(b) * ldo + u * 16 + f] = __float2bfloat16(o);
        } else {
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            const int col = (u * NG + q) * 16 + f;
            const float o = v[q] + epi_add[q];
            if (EPI == DP_EPI_F32) out_f32[static_cast<int64_t>(b) * ldo + col] = o;
            else out_bf[static_cast<int64_t>(b) * ldo + col] = __float2bfloat16(o);
          }
        }
      }
    }
    red_par ^= 1;  // the next unit writes the other buffer; its barrier orders this unit's reads before the reuse after it
  }
}

// ------------------------------------------------------------------------------------------------ attention phase
// M-RoPE (mcore_adapter/.../modeling_qwen2_5_vl.py:21-41) on q and the new k, KV-cache append, split-KV attention of the
// new token against the cache (the arithmetic of mrope_kv_kernel + decode_attn_kernel in sr_elementwise.cu / sr_tail.cu).
// Unit = (sample, kv head, key split), exactly one per CTA (n_units <= gridDim.x).  The phase is a chain of dependent global
// round trips (~1-2 us each while the producer saturates HBM), so it is arranged to need as few as possible:
//   * kv_start / kv_len / cache_row and the step's cos/sin table sit in shared memory (loaded once per launch);
//   * the K rows and V^T vectors of the OLD keys do not depend on this layer's qkv: they are requested into registers
//     BEFORE the grid barrier that ends the qkv phase (dp_attn_prefetch) and arrive while the CTA waits there;
//   * after the barrier: one round trip for q (+ the new k, v of the owning split, which never re-reads them from the
//     cache: they go to shared memory and, fire-and-forget, to the cache), then scores / softmax / P.V from registers;
//   * partial (o, m, l) -> workspace, one acq_rel ticket; the last CTA of a (sample, kv head) merges with ALL its loads
//     (statistics and partial outputs) in flight at once.
struct DpAttnUnit {
  int b, kvh, split, bk;
  int start, len, crow;
  int k_begin, k_end;
  bool active, owner;  // owner: this split contains the newest key (index len - 1 = cache row crow)
};
SR_DEVINL DpAttnUnit dp_attn_unit(const DpParams& p, const int* s_scal) {
  DpAttnUnit U;
  const int unit = blockIdx.x, ns = p.ns;
  U.active = unit < p.B * p.nkv * ns;
  U.split = unit % ns;
  U.bk = unit / ns;
  U.kvh = U.bk % p.nkv;
  U.b = U.active ? U.bk / p.nkv : 0;
  U.start = s_scal[U.b];
  U.len = s_scal[8 + U.b];
  U.crow = s_scal[16 + U.b];
  int per = (U.len + ns - 1) / ns;
  per = (per + 7) & ~7;  // V^T vectors are 8 keys wide
  U.k_begin = U.split * per;
  U.k_end = min(U.len, U.k_begin + per);
  U.owner = U.active && U.len > 0 && (U.len - 1) / per == U.split;
  return U;
}
// ------------------------------------------------------------------------------------------------ attention phase (mma)
// Compact formulation on mma.sync (the instruction footprint of a persistent kernel matters: everything beyond the 32 KB
// L1.5 instruction cache is re-fetched from L2 every layer, in competition with 6 TB/s of weight traffic).
// Scores: A = K rows (16 keys per warp, 8 warps = one 128-key pass), B = q (8 heads); P.V: A = V^T rows (16 output dims per
// warp), B = P (bf16, shared memory).  As in the GEMV phases a thread's 16-byte global load IS an A fragment (the k index is
// permuted identically in both operands), so every cache element is requested exactly once, with 16 loads per thread and pass
// that are issued BEFORE the grid barrier that ends the qkv phase (they do not depend on this layer's qkv).
// Thread (warp w, g = lane / 4, t = lane % 4) owns heads 2t, 2t+1 in both accumulators:
//   scores  c0..c3 = (key 16w+g | 16w+g+8) x (head 2t | 2t+1),   output c0..c3 = (dim 16w+g | 16w+g+8) x (head 2t | 2t+1).
// The newest key is never read from the cache: its roped k / v go to shared memory (and, fire-and-forget, to the cache) and
// enter the online softmax as one more key of the owning split.
constexpr int DPB_ROW = 320;                           // bytes per bf16 [128] row (+64 B: conflict-free 16 B fragment reads)
constexpr int DPB_Q = 0;                               // bf16 [8 heads][128]: roped q
constexpr int DPB_P = DPB_Q + 8 * DPB_ROW;             // bf16 [8 heads][128 keys]: probabilities of the pass
constexpr int DPB_RED = DPB_P + 8 * DPB_ROW;           // float [2][8 warps][8 heads]
constexpr int DPB_KNEW = DPB_RED + 2 * 8 * 8 * 4;      // float [128] roped k of the new token (bf16-rounded values)
constexpr int DPB_VNEW = DPB_KNEW + 128 * 4;           // float [128]
constexpr int DPB_SNEW = DPB_VNEW + 128 * 4;           // float [8] score of the new key per head (log2 domain)
constexpr int DPB_MS = DPB_SNEW + 8 * 4;               // merge: float m[32][8], l[32][8], f[32][8]
constexpr int DPB_LL = DPB_MS + 3 * DP_NS_MAX * 8 * 4; // float [8]
constexpr int DPB_FLAG = DPB_LL + 8 * 4;               // int
constexpr int DPB_END = DPB_FLAG + 16;
static_assert(DPB_END <= DP_XS_BYTES, "attention scratch must fit in the staged-activation region");

struct DpAttnFrag {
  uint4 k[8];  // k[c], k[4+c]: key rows 16w+g, 16w+g+8, dims 32c + 8t .. +8
  uint4 v[8];  // v[c], v[4+c]: V^T rows (dims) 16w+g, 16w+g+8, keys 32c + 8t .. +8 of the pass
};
// keys [p0, p0 + nk) of the split, all of them OLD keys (the caller clips the range at the newest key)
SR_DEVINL void dp_attn_load(const DpParams& p, const DpAttnUnit& U, const __nv_bfloat16* Kc, const __nv_bfloat16* Vt, int p0, int nk,
                             DpAttnFrag& f) {
  const int ct = threadIdx.x - 32, lane = ct & 31, w = ct >> 5, g = lane >> 2, t = lane & 3;
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = 16 * w + g + 8 * r;
    const __nv_bfloat16* kr = Kc + (static_cast<int64_t>(U.start + p0 + key) * p.nkv + U.kvh) * 128 + t * 8;
#pragma unroll
    for (int c = 0; c < 4; ++c) f.k[4 * r + c] = key < nk ? dp_ldcg(kr + c * 32) : z;
    const __nv_bfloat16* vr = Vt + static_cast<int64_t>(U.kvh * 128 + 16 * w + g + 8 * r) * p.ldvt + U.start + p0 + t * 8;
#pragma unroll
    for (int c = 0; c < 4; ++c) f.v[4 * r + c] = (c * 32 + t * 8) < nk ? dp_ldcg(vr + c * 32) : z;
  }
}
// zero the V^T elements of keys >= nk (stale cache columns must not meet a zero probability as NaN / Inf)
SR_DEVINL uint4 dp_mask_keys(uint4 v, int nvalid) {
  if (nvalid >= 8) return v;
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = w[j] & ((2 * j < nvalid ? 0xffffu : 0u) | (2 * j + 1 < nvalid ? 0xffff0000u : 0u));
  return make_uint4(o[0], o[1], o[2], o[3]);
}
SR_DEVINL int dp_attn_old_end(const DpAttnUnit& U) { return U.owner ? min(U.k_end, U.len - 1) : U.k_end; }
SR_DEVINL void dp_attn_prefetch(const DpParams& p, int layer, const DpAttnUnit& U, DpAttnFrag& f) {
  const int ke = dp_attn_old_end(U);
  if (!U.active || U.k_begin >= ke) return;
  dp_attn_load(p, U, p.Kc + static_cast<int64_t>(layer) * p.k_layer_stride, p.Vt + static_cast<int64_t>(layer) * p.vt_layer_stride,
                U.k_begin, min(DP_KEYS, ke - U.k_begin), f);
}

SR_DEVINL void dp_attention_phase(const DpParams& p, int layer, uint8_t* sm, const float* s_cos, const float* s_sin,
                                   const DpAttnUnit& U, DpAttnFrag& f) {
  if (!U.active) return;  // uniform per CTA
  const int ct = threadIdx.x - 32, lane = ct & 31, w = ct >> 5, g = lane >> 2, t = lane & 3;
  const int G = p.nq / p.nkv, ns = p.ns;
  float* s_red = reinterpret_cast<float*>(sm + DPB_RED);
  float* s_knew = reinterpret_cast<float*>(sm + DPB_KNEW);
  float* s_vnew = reinterpret_cast<float*>(sm + DPB_VNEW);
  float* s_snew = reinterpret_cast<float*>(sm + DPB_SNEW);
  float* m_ms = reinterpret_cast<float*>(sm + DPB_MS);
  float* m_ls = m_ms + DP_NS_MAX * 8;
  float* m_f = m_ls + DP_NS_MAX * 8;
  float* m_L = reinterpret_cast<float*>(sm + DPB_LL);
  int* s_flag = reinterpret_cast<int*>(sm + DPB_FLAG);
  __nv_bfloat16* Kc = p.Kc + static_cast<int64_t>(layer) * p.k_layer_stride;
  __nv_bfloat16* Vt = p.Vt + static_cast<int64_t>(layer) * p.vt_layer_stride;
  const int ldq = (p.nq + 2 * p.nkv) * 128;
  const int b = U.b, kvh = U.kvh;
  const __nv_bfloat16* qrow = p.qkv + static_cast<int64_t>(b) * ldq;
  const float* cs = s_cos + b * 64;
  const float* sn = s_sin + b * 64;
  // one round trip: q of the heads (threads 0..63: head ct/8, dims 8(ct%8).. and +64), new k (64..127), new v (128..255)
  if (ct < 64) {
    const int h = ct >> 3, d0 = (ct & 7) * 8;
    uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
    if (h < G) {
      const __nv_bfloat16* qh = qrow + (kvh * G + h) * 128;
      lo = dp_ldcg(qh + d0);
      hi = dp_ldcg(qh + 64 + d0);
    }
    const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // rotate-half; one rounding to bf16, like the q the per-layer path writes back
      const float a1 = bf16_lo(lw[j]), a2 = bf16_lo(hw[j]), b1 = bf16_hi(lw[j]), b2 = bf16_hi(hw[j]);
      const float ca = cs[d0 + 2 * j], sa = sn[d0 + 2 * j], cb = cs[d0 + 2 * j + 1], sb = sn[d0 + 2 * j + 1];
      o1[j] = pack_bf16x2(a1 * ca - a2 * sa, b1 * cb - b2 * sb);
      o2[j] = pack_bf16x2(a2 * ca + a1 * sa, b2 * cb + b1 * sb);
    }
    *reinterpret_cast<uint4*>(sm + DPB_Q + h * DPB_ROW + d0 * 2) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4*>(sm + DPB_Q + h * DPB_ROW + (64 + d0) * 2) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  } else if (U.owner && ct < 128) {
    const int d = ct - 64;
    const __nv_bfloat16* kh = qrow + (p.nq + kvh) * 128;
    const float x1 = __bfloat162float(__ldcg(kh + d)), x2 = __bfloat162float(__ldcg(kh + d + 64));
    const __nv_bfloat16 k1 = __float2bfloat16(x1 * cs[d] - x2 * sn[d]), k2 = __float2bfloat16(x2 * cs[d] + x1 * sn[d]);
    __nv_bfloat16* kd = Kc + (static_cast<int64_t>(U.crow) * p.nkv + kvh) * 128;
    kd[d] = k1;
    kd[d + 64] = k2;
    s_knew[d] = __bfloat162float(k1);
    s_knew[d + 64] = __bfloat162float(k2);
  } else if (U.owner && ct >= 128) {
    const int d = ct - 128;
    const __nv_bfloat16 v = __ldcg(qrow + (p.nq + p.nkv + kvh) * 128 + d);
    Vt[static_cast<int64_t>(kvh * 128 + d) * p.ldvt + U.crow] = v;
    s_vnew[d] = __bfloat162float(v);
  }
  dp_cbar();
  if (U.owner) {  // score of the new key: warp w <-> head w
    const uint2 qv = *reinterpret_cast<const uint2*>(sm + DPB_Q + w * DPB_ROW + lane * 8);
    const float4 kv = *reinterpret_cast<const float4*>(s_knew + lane * 4);
    float d = bf16_lo(qv.x) * kv.x + bf16_hi(qv.x) * kv.y + bf16_lo(qv.y) * kv.z + bf16_hi(qv.y) * kv.w;
    d = warp_sum(d);
    if (lane == 0) s_snew[w] = d * p.scale_log2;
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // heads 2t, 2t+1: running max, this thread's share of the sum
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  const int k_old_end = dp_attn_old_end(U);
  const uint8_t* qf = sm + DPB_Q + g * DPB_ROW + t * 16;
  const uint8_t* pf = sm + DPB_P + g * DPB_ROW + t * 16;
  int par = 0;
  for (int p0 = U.k_begin; p0 < k_old_end; p0 += DP_KEYS) {
    const int nk = min(DP_KEYS, k_old_end - p0);
    if (p0 != U.k_begin) dp_attn_load(p, U, Kc, Vt, p0, nk, f);  // first pass: requested before the grid barrier
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 xv = *reinterpret_cast<const uint4*>(qf + c * 64);
      dp_mma(s, f.k[c].x, f.k[4 + c].x, f.k[c].y, f.k[4 + c].y, xv.x, xv.y);
      dp_mma(s, f.k[c].z, f.k[4 + c].z, f.k[c].w, f.k[4 + c].w, xv.z, xv.w);
    }
    const bool ok0 = 16 * w + g < nk, ok1 = 16 * w + g + 8 < nk;
    s[0] = ok0 ? s[0] * p.scale_log2 : -INFINITY;
    s[1] = ok0 ? s[1] * p.scale_log2 : -INFINITY;
    s[2] = ok1 ? s[2] * p.scale_log2 : -INFINITY;
    s[3] = ok1 ? s[3] * p.scale_log2 : -INFINITY;
    float x0 = fmaxf(s[0], s[2]), x1 = fmaxf(s[1], s[3]);
#pragma unroll
    for (int d = 4; d < 32; d <<= 1) {
      x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, d));
      x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, d));
    }
    float* rp = s_red + par * 64;
    if (g == 0) *reinterpret_cast<float2*>(rp + w * 8 + 2 * t) = make_float2(x0, x1);
    dp_cbar();
#pragma unroll
    for (int w8 = 0; w8 < DP_CONSUMERS; ++w8) {
      const float2 v = *reinterpret_cast<const float2*>(rp + w8 * 8 + 2 * t);
      x0 = fmaxf(x0, v.x);
      x1 = fmaxf(x1, v.y);
    }
    const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);
    const float u0 = n0 == -INFINITY ? 0.f : n0, u1 = n1 == -INFINITY ? 0.f : n1;
    const float a0 = exp2f(m0 - u0), a1 = exp2f(m1 - u1);
    m0 = n0;
    m1 = n1;
    const float e0 = exp2f(s[0] - u0), e1 = exp2f(s[1] - u1), e2 = exp2f(s[2] - u0), e3 = exp2f(s[3] - u1);
    l0 = l0 * a0 + e0 + e2;
    l1 = l1 * a1 + e1 + e3;
    o[0] *= a0; o[1] *= a1; o[2] *= a0; o[3] *= a1;
    __nv_bfloat16* pr = reinterpret_cast<__nv_bfloat16*>(sm + DPB_P);
    const int kq = 16 * w + g;
    pr[(2 * t) * (DPB_ROW / 2) + kq] = __float2bfloat16(e0);
    pr[(2 * t + 1) * (DPB_ROW / 2) + kq] = __float2bfloat16(e1);
    pr[(2 * t) * (DPB_ROW / 2) + kq + 8] = __float2bfloat16(e2);
    pr[(2 * t + 1) * (DPB_ROW / 2) + kq + 8] = __float2bfloat16(e3);
    dp_cbar();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 pv = *reinterpret_cast<const uint4*>(pf + c * 64);
      const int nv = nk - (c * 32 + t * 8);
      const uint4 va = dp_mask_keys(f.v[c], nv), vb = dp_mask_keys(f.v[4 + c], nv);
      dp_mma(o, va.x, vb.x, va.y, vb.y, pv.x, pv.y);
      dp_mma(o, va.z, vb.z, va.w, vb.w, pv.z, pv.w);
    }
    par ^= 1;
  }
  if (U.owner) {  // the new key: one more online-softmax update (s_snew was written before the loop's barriers or the one below)
    if (U.k_begin >= k_old_end) dp_cbar();
    const float sn0 = s_snew[2 * t], sn1 = s_snew[2 * t + 1];
    const float n0 = fmaxf(m0, sn0), n1 = fmaxf(m1, sn1);
    const float a0 = exp2f(m0 - n0), a1 = exp2f(m1 - n1), e0 = exp2f(sn0 - n0), e1 = exp2f(sn1 - n1);
    m0 = n0;
    m1 = n1;
    const float v0 = s_vnew[16 * w + g], v1 = s_vnew[16 * w + g + 8];
    o[0] = o[0] * a0 + e0 * v0;
    o[1] = o[1] * a1 + e1 * v0;
    o[2] = o[2] * a0 + e0 * v1;
    o[3] = o[3] * a1 + e1 * v1;
    l0 = l0 * a0 + ((w == 0 && g == 0) ? e0 : 0.f);
    l1 = l1 * a1 + ((w == 0 && g == 0) ? e1 : 0.f);
  }
  // softmax denominators: sum this thread's shares over the keys (g, warps)
#pragma unroll
  for (int d = 4; d < 32; d <<= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, d);
    l1 += __shfl_xor_sync(0xffffffffu, l1, d);
  }
  dp_cbar();  // (the last pass's reads of s_red are done)
  if (g == 0) *reinterpret_cast<float2*>(s_red + w * 8 + 2 * t) = make_float2(l0, l1);
  // publish this split's un-normalised partial (o, m, l)
  float* wo = p.attn_ws + (static_cast<int64_t>(U.bk) * ns + U.split) * DP_PART;
  const int d0 = 16 * w + g;
  wo[(2 * t) * 128 + d0] = o[0];
  wo[(2 * t + 1) * 128 + d0] = o[1];
  wo[(2 * t) * 128 + d0 + 8] = o[2];
  wo[(2 * t + 1) * 128 + d0 + 8] = o[3];
  dp_cbar();
  if (ct < 4) {  // lanes t = 0..3 of warp 0 hold the running maxima of heads 2t, 2t+1
    float L0 = 0.f, L1 = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < DP_CONSUMERS; ++w8) {
      L0 += s_red[w8 * 8 + 2 * ct];
      L1 += s_red[w8 * 8 + 2 * ct + 1];
    }
    wo[DP_GMAX * 128 + 2 * ct] = m0;
    wo[DP_GMAX * 128 + 2 * ct + 1] = m1;
    wo[DP_GMAX * 128 + DP_GMAX + 2 * ct] = L0;
    wo[DP_GMAX * 128 + DP_GMAX + 2 * ct + 1] = L1;
  }
  dp_cbar();
  if (ct == 0) {  // release our partial (cumulative through the CTA barrier), acquire the others'
    int ticket;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(ticket) : "l"(p.attn_tickets + U.bk) : "memory");
    *s_flag = (ticket == ns - 1) ? 1 : 0;
  }
  dp_cbar();
  if (*s_flag == 0) return;
  // merge (last CTA of this (sample, kv head)): statistics and partial outputs requested together
  const float* wb = p.attn_ws + static_cast<int64_t>(U.bk) * ns * DP_PART;
  float ov[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* src = wb + (2 * t + (i & 1)) * 128 + d0 + 8 * (i >> 1);
#pragma unroll
    for (int j = 0; j < 16; ++j) ov[i][j] = j < ns ? __ldcg(src + j * DP_PART) : 0.f;
  }
  for (int i = ct; i < ns * 8; i += DP_CTHREADS) {
    const int s2 = i >> 3, h = i & 7;
    m_ms[i] = __ldcg(wb + s2 * DP_PART + DP_GMAX * 128 + h);
    m_ls[i] = __ldcg(wb + s2 * DP_PART + DP_GMAX * 128 + DP_GMAX + h);
  }
  dp_cbar();
  if (ct < 8) {
    float Mx = -INFINITY;
    for (int s2 = 0; s2 < ns; ++s2) Mx = fmaxf(Mx, m_ms[s2 * 8 + ct]);
    const float Mu = (Mx == -INFINITY) ? 0.f : Mx;
    float L = 0.f;
    for (int s2 = 0; s2 < ns; ++s2) {
      const float fz = exp2f(m_ms[s2 * 8 + ct] - Mu);
      m_f[s2 * 8 + ct] = fz;
      L += m_ls[s2 * 8 + ct] * fz;
    }
    m_L[ct] = L;
  }
  dp_cbar();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int h = 2 * t + (i & 1);
    float O = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) O += j < ns ? ov[i][j] * m_f[j * 8 + h] : 0.f;
    for (int j = 16; j < ns; ++j) O += __ldcg(wb + j * DP_PART + h * 128 + d0 + 8 * (i >> 1)) * m_f[j * 8 + h];
    const float L = m_L[h];
    if (h < G) p.ao[static_cast<int64_t>(b) * (p.nq * 128) + (kvh * G + h) * 128 + d0 + 8 * (i >> 1)] = __float2bfloat16(L > 0.f ? O / L : 0.f);
  }
  if (ct == 0) p.attn_tickets[U.bk] = 0;  // re-arm for the next layer / launch
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(DP_THREADS, 1) decode_persistent_kernel(const DpParams p) {
  extern __shared__ uint8_t dp_smem_raw[];
  uint8_t* smem = dp_smem_raw + ((1024u - (smem_u32(dp_smem_raw) & 1023u)) & 1023u);
  DpRing ring;
  ring.data = smem + DP_OFF_RING;
  ring.full = reinterpret_cast<uint64_t*>(smem + DP_OFF_BAR);
  ring.empty = ring.full + DP_STAGES;
  ring.slot = 0;
  ring.phase = 0;
  ring.n = p.ring_stages;
  uint8_t* xs = smem + DP_OFF_XS;
  float* red = reinterpret_cast<float*>(smem + DP_OFF_RED);

  DpLayer* s_layers = reinterpret_cast<DpLayer*>(smem + DP_OFF_LAYERS);
  float* s_cos = reinterpret_cast<float*>(smem + DP_OFF_ROPE);
  float* s_sin = s_cos + 8 * 64;
  int* s_scal = reinterpret_cast<int*>(smem + DP_OFF_SCAL);
  volatile int* s_pcnt = reinterpret_cast<volatile int*>(smem + DP_OFF_PCNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // per-launch constants -> shared memory (pointer / scalar loads off the critical paths)
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.layers);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_layers);
    for (int i = threadIdx.x; i < p.n_layers * 7; i += DP_THREADS) dst[i] = src[i];
    for (int i = threadIdx.x; i < 8 * 64; i += DP_THREADS) {
      const bool ok = i < p.B * 64;
      s_cos[i] = ok ? p.cosb[i] : 0.f;
      s_sin[i] = ok ? p.sinb[i] : 0.f;
    }
    if (threadIdx.x < 8) {
      const bool ok = static_cast<int>(threadIdx.x) < p.B;
      s_scal[threadIdx.x] = ok ? p.kv_start[threadIdx.x] : 0;
      s_scal[8 + threadIdx.x] = ok ? p.kv_len[threadIdx.x] : 0;
      s_scal[16 + threadIdx.x] = ok ? p.cache_row[threadIdx.x] : 0;
    }
  }
  if (threadIdx.x == 0) {
    *s_pcnt = 0;
#pragma unroll
    for (int i = 0; i < DP_STAGES; ++i) {
      mbar_init(&ring.full[i], 1);
      mbar_init(&ring.empty[i], DP_CONSUMERS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int kq = DP_H >> 5;                       // k-chunks of the hidden-size matrices
  const int n_qkv = (p.nq + 2 * p.nkv) * 128 / 16;  // 16-row units
  const int n_o = DP_H / 16;
  const int n_gu = p.I / 16;                      // units of 32 packed rows (gate|up pairs) = I/16
  const int kd = p.I >> 5;
  const int n_head = p.vocab / 16;
  const int limit = p.stop_after > 0 ? p.stop_after : 0x7fffffff;
  // Unit -> CTA mapping, rotated per matrix: unit u of a matrix goes to CTA (u + offset) mod gridDim.  The unit counts (80 / 128
  // / 688 / 128 on 148 CTAs) leave some CTAs without a unit, or with one unit more; the offsets make those sets complement each
  // other (e.g. the 20 CTAs idle in the down projection take a fifth gate/up unit), which lowers the bytes of the busiest CTA.
  auto vidx = [&](int off) { const int g = static_cast<int>(gridDim.x); return (static_cast<int>(blockIdx.x) + g - (off % g)) % g; };
  const int vq = vidx(p.unit_off[0]), vo = vidx(p.unit_off[1]), vg = vidx(p.unit_off[2]), vd = vidx(p.unit_off[3]);

  if (warp == 0) {
    // ---------------- producer: weights only; runs ahead of the consumers across phases ----------------
    if (lane == 0) {
      uint64_t policy = 0;
      const bool hint = (p.flags & 1) != 0;
      if (hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      int ph = 0, issued = 0;
      dp_trace_at(p, 0, 0);
      for (int l = 0; l < p.n_layers; ++l) {
        const DpLayer L = s_layers[l];
        if (ph++ >= limit) return;
        dp_produce<2, false>(ring, vq, L.qkv_wp, n_qkv / 2, kq, policy, hint, s_pcnt, issued);
        dp_trace_at(p, 0, 1 + 4 * l);
        if (ph++ >= limit) return;  // attention phase: no weights
        if (ph++ >= limit) return;
        dp_produce<1, false>(ring, vo, L.o_wp, n_o, kq, policy, hint, s_pcnt, issued);
        dp_trace_at(p, 0, 2 + 4 * l);
        if (ph++ >= limit) return;
        dp_produce<2, false>(ring, vg, L.gu_wp, n_gu, kq, policy, hint, s_pcnt, issued);
        dp_trace_at(p, 0, 3 + 4 * l);
        if (ph++ >= limit) return;
        dp_produce<1, true>(ring, vd, L.down_wp, n_o, kd, policy, hint, s_pcnt, issued);
        dp_trace_at(p, 0, 4 + 4 * l);
      }
      if (ph++ >= limit) return;
      dp_produce<1, false>(ring, static_cast<int>(blockIdx.x), p.lm_head_p, n_head, kq, policy, hint, s_pcnt, issued);
      dp_trace_at(p, 0, 1 + 4 * p.n_layers);
    }
    return;
  }
  if (warp == DP_CONSUMERS + 1) {
    // ---------------- L2 prefetcher: same walk as the producer, `pf_stages` ahead of it ----------------
    if (lane == 0 && p.pf_stages > 0) {
      int ph = 0, sidx = 0;
      const int lead = p.pf_stages;
      uint64_t pf_policy = 0;
      const bool pf_hint = (p.flags & 128) != 0;
      if (pf_hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf_policy));
      for (int l = 0; l < p.n_layers; ++l) {
        const DpLayer L = s_layers[l];
        if (ph++ >= limit) return;
        dp_prefetch_l2<2, false>(sidx, s_pcnt, lead, vq, L.qkv_wp, n_qkv / 2, kq, pf_policy, pf_hint);
        if (ph++ >= limit) return;
        if (ph++ >= limit) return;
        dp_prefetch_l2<1, false>(sidx, s_pcnt, lead, vo, L.o_wp, n_o, kq, pf_policy, pf_hint);
        if (ph++ >= limit) return;
        dp_prefetch_l2<2, false>(sidx, s_pcnt, lead, vg, L.gu_wp, n_gu, kq, pf_policy, pf_hint);
        if (ph++ >= limit) return;
        dp_prefetch_l2<1, true>(sidx, s_pcnt, lead, vd, L.down_wp, n_o, kd, pf_policy, pf_hint);
      }
      if (ph++ >= limit) return;
      dp_prefetch_l2<1, false>(sidx, s_pcnt, lead, static_cast<int>(blockIdx.x), p.lm_head_p, n_head, kq, pf_policy, pf_hint);
    }
    return;
  }

  // ---------------- consumers ----------------
  const int cw = warp - 1;
  int red_par = 0;
  unsigned int bar_gen = 0;
  int ph = 0, tr = 0;
  const unsigned int G = gridDim.x;
  const bool tracer = threadIdx.x == 32;
  const DpAttnUnit U = dp_attn_unit(p, s_scal);
  DpAttnFrag fr;
#define DP_TRACE() do { if (tracer) dp_trace(p, 1, tr); } while (0)
  DP_TRACE();
  for (int l = 0; l < p.n_layers; ++l) {
    const DpLayer L = s_layers[l];
    // A: qkv = rmsnorm(x; ln1) . Wqkv^T + b
    if (ph++ >= limit) return;
    dp_stage_rows<true>(xs, p.x, p.ldx, p.B, L.ln1, p.eps, cw, lane);
    dp_cbar();
    // units of 32 rows: (nq + 2 nkv) * 8 = 160 16-row groups would leave 12 of 148 CTAs with two serial units; 80 double
    // units fit the ring (128 KB) and take one pass (the 68 CTAs without a unit let their producers run ahead)
    dp_gemv_phase<2, DP_EPI_BIAS, true>(ring, vq, n_qkv / 2, kq, xs, nullptr, 0, p.B, red, red_par, L.qkv_b, p.qkv, nullptr,
                                        static_cast<int64_t>(n_qkv) * 16, cw, lane);
    const bool do_attn = ph < limit;
    if (do_attn) dp_attn_prefetch(p, l, U, fr);  // old K / V^T rows fly while the CTA waits at the barrier
    DP_TRACE();
    dp_grid_barrier(p.grid_bar, ++bar_gen * G);
    DP_TRACE();
    // B: rope + KV append + attention -> ao
    if (ph++ >= limit) return;
    dp_attention_phase(p, l, xs, s_cos, s_sin, U, fr);
    DP_TRACE();
    dp_grid_barrier(p.grid_bar, ++bar_gen * G);
    DP_TRACE();
    // C: x += ao . Wo^T
    if (ph++ >= limit) return;
    dp_stage_rows<false>(xs, p.ao, DP_H, p.B, nullptr, 0.f, cw, lane);
    dp_cbar();
    dp_gemv_phase<1, DP_EPI_RESIDUAL, true>(ring, vo, n_o, kq, xs, nullptr, 0, p.B, red, red_par, nullptr, p.x, nullptr, p.ldx, cw, lane);
    DP_TRACE();
    dp_grid_barrier(p.grid_bar, ++bar_gen * G);
    DP_TRACE();
    // D: act = silu(gate) * up, [gate|up] = rmsnorm(x; ln2) . Wgu^T
    if (ph++ >= limit) return;
    dp_stage_rows<true>(xs, p.x, p.ldx, p.B, L.ln2, p.eps, cw, lane);
    dp_cbar();
    dp_gemv_phase<2, DP_EPI_SWIGLU, true>(ring, vg, n_gu, kq, xs, nullptr, 0, p.B, red, red_par, nullptr, p.act, nullptr, p.I, cw, lane);
    DP_TRACE();
    dp_grid_barrier(p.grid_bar, ++bar_gen * G);
    DP_TRACE();
    // E: x += act . Wdown^T
    if (ph++ >= limit) return;
    dp_gemv_phase<1, DP_EPI_RESIDUAL, false>(ring, vd, n_o, kd, nullptr, p.act, p.I, p.B, red, red_par, nullptr, p.x, nullptr, p.ldx, cw, lane);
    DP_TRACE();
    dp_grid_barrier(p.grid_bar, ++bar_gen * G);
    DP_TRACE();
  }
  // final norm + LM head -> fp32 logits
  if (ph++ >= limit) return;
  dp_stage_rows<true>(xs, p.x, p.ldx, p.B, p.final_norm, p.eps, cw, lane);
  dp_cbar();
  dp_gemv_phase<1, DP_EPI_F32, true>(ring, static_cast<int>(blockIdx.x), n_head, kq, xs, nullptr, 0, p.B, red, red_par, nullptr, nullptr, p.logits, p.ld_logits, cw, lane);
  DP_TRACE();
#undef DP_TRACE
}

}  // namespace sr

using namespace sr;

extern "C" int64_t sr_decode_persistent_workspace_bytes(int batch, int n_kv_heads) {
  // [16 B: grid barrier counter][tickets: batch*nkv int32, 16 B aligned][partials: batch*nkv*DP_NS_MAX*DP_PART fp32]
  const int64_t bk = static_cast<int64_t>(batch) * n_kv_heads;
  return 16 + ((bk * 4 + 15) / 16) * 16 + bk * DP_NS_MAX * DP_PART * 4;
}

static unsigned long long* g_dp_trace = nullptr;
static int g_dp_trace_events = 0;
extern "C" int sr_set_decode_persistent_trace(void* buffer, int events_per_role) {
  g_dp_trace = reinterpret_cast<unsigned long long*>(buffer);
  g_dp_trace_events = buffer ? events_per_role : 0;
  return SR_OK;
}

static int g_dp_pf_stages = -1;  // -1: env SR_DECODE_PF_STAGES (default below)
static int g_dp_flags = -1;      // -1: env SR_DECODE_FLAGS
static int g_dp_ring = 0;        // 0: env SR_DECODE_RING (default DP_STAGES); set through bits 8.. of the tuning flags
extern "C" int sr_set_decode_persistent_tuning(int pf_stages, int flags) {
  if (pf_stages < 0 || pf_stages > 4096) return SR_ERR_INVALID;
  g_dp_pf_stages = pf_stages;
  g_dp_flags = flags & 0xff;
  g_dp_ring = (flags >> 8) & 0xff;  // tuning aid: ring stages in bits 8..15 (0 = default)
  if (g_dp_ring == 1 || g_dp_ring > DP_STAGES) return SR_ERR_INVALID;
  return SR_OK;
}

static int g_dp_unit_off[4] = {-1, 0, 0, 0};  // -1: env SR_DECODE_UNIT_OFFSETS="q,o,g,d" (default below)
extern "C" int sr_set_decode_persistent_offsets(int qkv, int o, int gate_up, int down) {
  if (qkv < 0 || o < 0 || gate_up < 0 || down < 0) return SR_ERR_INVALID;
  g_dp_unit_off[0] = qkv;
  g_dp_unit_off[1] = o;
  g_dp_unit_off[2] = gate_up;
  g_dp_unit_off[3] = down;
  return SR_OK;
}

static int g_dp_coop = -1;
extern "C" int sr_set_decode_persistent_coop(int on) {
  g_dp_coop = on ? 1 : 0;
  return SR_OK;
}

extern "C" int sr_decode_step_persistent(const void* layer_table, int n_layers, const void* lm_head_packed, const float* final_norm,
                                         int vocab, void* x, int64_t ldx, void* qkv_scratch, void* ao_scratch, void* act_scratch,
                                         float* logits, int64_t ld_logits, const float* cosb, const float* sinb,
                                         const int32_t* cache_row, const int32_t* kv_start, const int32_t* kv_len, void* Kc,
                                         int64_t k_layer_stride, void* Vt, int64_t vt_layer_stride, int64_t ldvt, int batch,
                                         int n_q_heads, int n_kv_heads, int intermediate, float norm_eps, float scale,
                                         void* workspace, int stop_after_phases, void* stream) {
  SR_REQUIRE(layer_table && lm_head_packed && final_norm && x && qkv_scratch && ao_scratch && act_scratch && logits && cosb && sinb &&
                 cache_row && kv_start && kv_len && Kc && Vt && workspace,
             "sr_decode_step_persistent: null pointer");
  SR_REQUIRE(batch > 0 && batch <= 8, "sr_decode_step_persistent: batch must be in [1, 8] (got %d)", batch);
  SR_REQUIRE(n_q_heads * 128 == DP_H && n_kv_heads > 0 && n_q_heads % n_kv_heads == 0 && n_q_heads / n_kv_heads <= DP_GMAX,
             "sr_decode_step_persistent: needs hidden 2048 = n_q_heads * 128 and <= 8 query heads per kv head");
  SR_REQUIRE(n_layers <= DP_MAX_LAYERS, "sr_decode_step_persistent: at most %d layers", DP_MAX_LAYERS);
  SR_REQUIRE(intermediate % 32 == 0 && vocab % 16 == 0 && n_layers > 0, "sr_decode_step_persistent: intermediate %% 32, vocab %% 16");
  SR_REQUIRE(ldx % 8 == 0 && ldvt % 8 == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(act_scratch) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(ao_scratch) & 15) == 0 && (reinterpret_cast<uintptr_t>(lm_head_packed) & 15) == 0,
             "sr_decode_step_persistent: alignment");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static int grid_cached = 0;
  if (grid_cached == 0) {
    SR_CHECK_CUDA(cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_BYTES));
    int per_sm = 0;
    SR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_persistent_kernel, DP_THREADS, DP_SMEM_BYTES));
    SR_REQUIRE(per_sm >= 1, "sr_decode_step_persistent: kernel does not fit on an SM (smem %d B)", DP_SMEM_BYTES);
    grid_cached = sm_count();  // one CTA per SM: all co-resident (required by the grid barrier)
  }
  if (g_dp_coop < 0) {
    const char* e = getenv("SR_DECODE_COOP");
    g_dp_coop = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  DpParams p;
  p.layers = reinterpret_cast<const DpLayer*>(layer_table);
  p.n_layers = n_layers;
  p.lm_head_p = reinterpret_cast<const __nv_bfloat16*>(lm_head_packed);
  p.final_norm = final_norm;
  p.vocab = vocab;
  p.x = reinterpret_cast<__nv_bfloat16*>(x);
  p.ldx = ldx;
  p.qkv = reinterpret_cast<__nv_bfloat16*>(qkv_scratch);
  p.ao = reinterpret_cast<__nv_bfloat16*>(ao_scratch);
  p.act = reinterpret_cast<__nv_bfloat16*>(act_scratch);
  p.logits = logits;
  p.ld_logits = ld_logits;
  p.cosb = cosb;
  p.sinb = sinb;
  p.cache_row = cache_row;
  p.kv_start = kv_start;
  p.kv_len = kv_len;
  p.Kc = reinterpret_cast<__nv_bfloat16*>(Kc);
  p.k_layer_stride = k_layer_stride;
  p.Vt = reinterpret_cast<__nv_bfloat16*>(Vt);
  p.vt_layer_stride = vt_layer_stride;
  p.ldvt = ldvt;
  p.B = batch;
  p.nq = n_q_heads;
  p.nkv = n_kv_heads;
  p.I = intermediate;
  p.eps = norm_eps;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int64_t bk = static_cast<int64_t>(batch) * n_kv_heads;
  char* ws = reinterpret_cast<char*>(workspace);
  p.grid_bar = reinterpret_cast<unsigned int*>(ws);
  p.attn_tickets = reinterpret_cast<int*>(ws + 16);
  p.attn_ws = reinterpret_cast<float*>(ws + 16 + ((bk * 4 + 15) / 16) * 16);
  SR_REQUIRE(bk <= grid_cached, "sr_decode_step_persistent: batch * n_kv_heads must not exceed the SM count");
  int ns = grid_cached / static_cast<int>(bk);
  p.ns = ns < 1 ? 1 : (ns > DP_NS_MAX ? DP_NS_MAX : ns);
  p.stop_after = stop_after_phases;
  if (g_dp_pf_stages < 0) {
    const char* e = getenv("SR_DECODE_PF_STAGES");
    g_dp_pf_stages = e ? atoi(e) : 0;
  }
  if (g_dp_flags < 0) {
    const char* e = getenv("SR_DECODE_FLAGS");
    g_dp_flags = e ? atoi(e) : 1;  // evict_first on the weight stream: keeps activations AND the kernel's instructions in L2
  }
  static int ring_stages = -1;
  if (ring_stages < 0) {
    const char* e = getenv("SR_DECODE_RING");
    ring_stages = e ? atoi(e) : 8;  // 8 of the 10 stages measured best (2.13 vs 2.16 ms/step; LM head 92 vs 103 us)
    if (ring_stages < 2 || ring_stages > DP_STAGES) ring_stages = DP_STAGES;
  }
  p.ring_stages = g_dp_ring > 0 ? g_dp_ring : ring_stages;
  if (g_dp_unit_off[0] < 0) {
    const int dflt[4] = {96, 40, 0, 20};
    for (int i = 0; i < 4; ++i) g_dp_unit_off[i] = dflt[i];
    const char* e = getenv("SR_DECODE_UNIT_OFFSETS");
    if (e) sscanf(e, "%d,%d,%d,%d", &g_dp_unit_off[0], &g_dp_unit_off[1], &g_dp_unit_off[2], &g_dp_unit_off[3]);
  }
  for (int i = 0; i < 4; ++i) p.unit_off[i] = g_dp_unit_off[i] < 0 ? 0 : g_dp_unit_off[i];
  p.pf_stages = g_dp_pf_stages;
  p.flags = g_dp_flags;
  p.trace = g_dp_trace;
  p.trace_events = g_dp_trace_events;
  // the barrier counter (and, defensively, the tickets) start at zero for every launch
  SR_CHECK_CUDA(cudaMemsetAsync(ws, 0, 16 + ((bk * 4 + 15) / 16) * 16, s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_cached);
  cfg.blockDim = dim3(DP_THREADS);
  cfg.dynamicSmemBytes = DP_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dp_coop ? 1 : 0;
  SR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_persistent_kernel, p));
  return SR_OK;
}
