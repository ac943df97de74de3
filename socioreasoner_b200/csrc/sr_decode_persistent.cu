// sr_decode_persistent.cu -- one decode step (all decoder layers + final norm + LM head) as ONE persistent kernel.
//
// What it replaces: the per-token body of the reference's rollout decode loop (vLLM engine step,
// roll/distributed/strategy/vllm_strategy.py:114-141; HF eager equivalent: Qwen2_5_VLDecoderLayer x L with use_cache,
// HF:762-827, 672-759, 611-624, final norm + lm_head HF:1385) -- here for M <= 8 live sequences.
//
// Why: at M = 8 the step is pure weight streaming (6.2 GB for the 3B model).  Launched as ~6 kernels per layer, every kernel
// pays a ramp and a tail -- measured 2.84 ms/token against ~1.0 ms of streaming at the measured HBM rate.
// Here the HBM stream never stops:
//   * grid = one CTA per SM, resident for the whole step;
//   * warp 0 of every CTA is a PRODUCER: one thread walks the static schedule (layer -> {qkv, o, gate/up, down} -> this CTA's
//     row groups -> 16 KB stages) and issues cp.async.bulk copies of the pre-packed weights ([N/16][K/32][16][32], 1 KB
//     blocks) into a shared-memory ring guarded by full/empty mbarriers.  Weights never depend on activations, so the producer
//     runs AHEAD of the consumers across phase boundaries;
//   * warps 1-8 are CONSUMERS: per phase they stage the 8 activation rows in shared memory (RMSNorm fused in the staging,
//     HF:57-71), read 16-byte weight fragments from the ring straight into mma.sync.m16n8k16 A operands (same k permutation
//     as sr_gemv.cu: a 16 B load IS a fragment), reduce across warps, run the fused epilogue (bias / residual / SwiGLU);
//   * phases per layer: A qkv | B rope + KV append + split-KV attention (flash-decoding) + merge | C o-proj + residual |
//     D gate/up + SwiGLU | E down + residual; then final norm + LM head (fp32 logits).
//
// Round 2: NO grid barriers.  Round 1 separated the phases by 5 grid barriers per layer (release-add + acquire-poll + the
// dependent activation load after it = 3-4 serialized L2 round trips each, 21 us of the 58 us a layer took, against 24 us of
// weight streaming).  Now every activation that crosses CTAs (x after o-proj / after down, qkv, attention partials, attention
// output, the down projection's partial sums) travels in FLAG-CARRYING lines, the "LL" protocol NCCL uses for latency-bound
// collectives: a 16-byte line = {data0, gen, data1, gen}; the writer's store is fire-and-forget, the reader polls THE DATA ITSELF
// until both generation words match (each 8-byte half is written atomically and carries its own flag, so no fence, no counter, no
// acquire is needed) -- one L2 round trip per exchange instead of three or four, and a CTA only ever waits for the lines it
// actually reads.  gen = epoch + 1 + layer; the epoch lives in the workspace and is advanced by the last CTA to leave the kernel,
// so stale lines (previous layer / launch / CUDA-graph replay) never match.  Buffers are rewritten once per layer; a writer of
// layer l+1 has (transitively, through the data it polled) observed results of every reader of the layer-l version.
// The SwiGLU output does not cross CTAs at all: the CTA that computes 16 gate/up columns multiplies them by the matching
// 2048 x 16 K-slice of the down projection right away (dp_de_phase) and only the fp32 partial of x is exchanged and reduced in a
// fixed order (dp_reduce_x).
// Measured and NOT adopted: L2 prefetching of the weight stream during the exchange chain (an L2 hit costs
// the same L2 throughput as a miss's fill + read and the stream already runs at that limit); a TMEM-backed weight ring with banker
// warps (16 more stages of look-ahead per SM: 2 % faster stand-alone, 4 % slower inside the power-capped bench step).
// HBM-bound: algorithmic bytes per launch = all decoder weights + LM head (N*K*2 each) + KV cache of the live sequences.
#include <stdio.h>
#include <stdlib.h>

#include "sr_host.h"

namespace sr {

constexpr int DP_CONSUMERS = 8;
constexpr int DP_CTHREADS = DP_CONSUMERS * 32;
constexpr int DP_THREADS = DP_CTHREADS + 32;          // warp 0 = producer, warps 1-8 = consumers
constexpr int DP_STAGE_BLOCKS = 16;                   // 1 KB weight blocks per ring stage
constexpr int DP_STAGE_BYTES = DP_STAGE_BLOCKS * 1024;
constexpr int DP_STAGES = 10;
constexpr int DP_H = 2048;                            // staged activation width (hidden size = n_q_heads * 128)
constexpr int DP_HL = DP_H / 4;                       // LL lines per hidden row (4 bf16 per line)
constexpr int DP_XS_STRIDE = DP_H * 2 + 64;           // bytes per staged row (+64 B: conflict-free 16 B fragment reads)
constexpr int DP_MAX_LAYERS = 64;                     // layer table staged in shared memory
constexpr int DP_GMAX = 8;                            // query heads per kv head (max)
constexpr int DP_KEYS = 128;                          // keys per attention pass
constexpr int DP_NS_MAX = 32;                         // max key splits
constexpr int DP_PART_LINES = 128 * 4 + DP_GMAX;      // LL lines per attention partial: o[dim][head pair], then {m, l}[head]
constexpr int DP_MERGE_CHUNK = 10;                    // splits merged per round trip

// shared-memory carve-up (bytes, from a 1024-aligned base)
constexpr int DP_OFF_RING = 0;
constexpr int DP_OFF_XS = DP_STAGES * DP_STAGE_BYTES;                 // staged activations; aliased by attention scratch
constexpr int DP_XS_BYTES = 8 * DP_XS_STRIDE;                         // 33280
constexpr int DP_OFF_RED = DP_OFF_XS + DP_XS_BYTES;                   // [2][8 warps][2 rg][8 batch][20] fp32
constexpr int DP_RED_WQ = 8 * 20;                                     // floats per (warp, row group): [batch][16 features + 4 pad]
constexpr int DP_RED_BYTES = 2 * DP_CONSUMERS * 2 * DP_RED_WQ * 4;    // 20480
constexpr int DP_OFF_BAR = DP_OFF_RED + DP_RED_BYTES;
constexpr int DP_OFF_LAYERS = DP_OFF_BAR + 2 * DP_STAGES * 8 + 32;    // DpLayer[DP_MAX_LAYERS]
constexpr int DP_OFF_ROPE = DP_OFF_LAYERS + DP_MAX_LAYERS * 64;       // cos[8][64], sin[8][64] fp32 of this step
constexpr int DP_OFF_SCAL = DP_OFF_ROPE + 2 * 8 * 64 * 4;             // kv_start[8], kv_len[8], cache_row[8], epoch, gate, pcnt
constexpr int DP_OFF_ACT = DP_OFF_SCAL + 4 * 8 * 4;                   // bf16 [8 batch][16 k]: SwiGLU output of the current unit
constexpr int DP_OFF_RED4 = DP_OFF_ACT + 256;                         // float [4][64][2] (partial-sum reduction) + float [8][16]
constexpr int DP_SMEM_BYTES = DP_OFF_RED4 + 4 * 64 * 2 * 4 + 8 * 16 * 4 + 1024;  // + alignment slack
static_assert(DP_SMEM_BYTES <= 232448, "shared-memory budget of one sm_100 CTA");

struct DpLayer {  // == sr_decode_layer (include/socioreasoner_b200.h)
  const __nv_bfloat16* qkv_wp;
  const float* qkv_b;
  const float* ln1;
  const __nv_bfloat16* o_wp;
  const float* ln2;
  const __nv_bfloat16* gu_wp;
  const __nv_bfloat16* down_wk;  // K-sliced: [I/16 units][4 stages][8 warps][4][16 rows][16 k (quad order)] (weights.py::_kslice_pack)
  const void* reserved;
};

struct DpParams {
  const DpLayer* layers;
  int n_layers;
  const __nv_bfloat16* lm_head_p;
  const float* final_norm;
  int vocab;
  __nv_bfloat16* x;        // in: embedded tokens [B, 2048] (ldx); debug: hidden state dumps
  int64_t ldx;
  __nv_bfloat16* qkv;      // debug dumps (flags bit 1): [B, (nq + 2 nkv) * 128], [B, nq * 128]
  __nv_bfloat16* ao;
  __nv_bfloat16* act;      // debug dump of the SwiGLU output [B, I] (it never leaves the SMs otherwise)
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
  unsigned int* epoch;     // workspace header: [0] epoch, [1] CTAs that left the kernel
  uint4* xa_ll;            // LL buffers (16-byte lines {d0, gen, d1, gen}; 4 bf16 or 2 fp32 per line)
  uint4* xb_ll;
  uint4* qkv_ll;
  uint4* ao_ll;
  uint4* part_ll;          // [B*nkv][ns][DP_PART_LINES]
  uint4* xpart_ll;         // [gridDim.x][2048 rows * 4]: every CTA's fp32 partial of the down projection, line (row, batch pair)
  int ns;
  int stop_after;          // debug: run only the first `stop_after` phases (<= 0: all)
  int ring_stages;         // shared-memory ring stages in use (2 .. DP_STAGES)
  int unit_off[4];         // unit -> CTA rotation of the qkv / o / gate-up / down matrices
  int flags;               // bit 0: L2 evict_first hint on the ring copies; bit 1: dump plain intermediates (bring-up)
  unsigned long long* trace;  // optional [gridDim.x][2 roles][trace_events][2] (globaltimer ns, clock64)
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

// ------------------------------------------------------------------------------------------------ LL lines
#ifndef SR_LL_SPIN_LIMIT
#define SR_LL_SPIN_LIMIT (1u << 22)
#endif
SR_DEVINL void ll_store(uint4* p, uint32_t d0, uint32_t d1, uint32_t gen) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(d0), "r"(gen), "r"(d1), "r"(gen) : "memory");
}
// Optional back-off between polling rounds (compile-time experiment: -DSR_LL_BACKOFF=<ns>); off by default.
#ifdef SR_LL_BACKOFF
#define SR_LL_BACKOFF_SLEEP() __nanosleep(SR_LL_BACKOFF)
#else
#define SR_LL_BACKOFF_SLEEP() ((void)0)
#endif
SR_DEVINL uint4 ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
SR_DEVINL bool ll_ok(const uint4& v, uint32_t gen) { return v.y == gen && v.w == gen; }
// Re-poll one line until both halves carry `gen`.  The fast path (line already there) is two compares; the polling loop is a
// real function call so that the ~150 inlined wait sites do not bloat the instruction footprint of the persistent kernel.
// Bounded: a broken schedule traps (CUDA error through the C ABI) instead of hanging the GPU.
__device__ __noinline__ uint4 ll_spin_slow(const uint4* p, uint32_t gen) {
  unsigned int spins = 0;
  uint4 v = ll_load(p);
  while (!ll_ok(v, gen)) {
    SR_LL_BACKOFF_SLEEP();
    v = ll_load(p);
    if (++spins > SR_LL_SPIN_LIMIT) {
      printf("sr: LL wait timeout block=%d thread=%d want=%u have=%u/%u\n", blockIdx.x, threadIdx.x, gen, v.y, v.w);
      __trap();
    }
  }
  return v;
}
SR_DEVINL void ll_spin(uint4& v, const uint4* p, uint32_t gen) {
  if (!ll_ok(v, gen)) v = ll_spin_slow(p, gen);
}
__device__ __noinline__ void ll_timeout(uint32_t gen) {
  printf("sr: LL wait timeout block=%d thread=%d want=%u\n", blockIdx.x, threadIdx.x, gen);
  __trap();
}
// Wait for N lines that were all requested already (v[i] = ll_load(addr(i))).  Lines that had not been written yet are NOT
// polled one after the other (that would serialise up to N L2 round trips): every round re-requests ALL missing lines at once,
// so the thread has its data one round trip after the last line was written.
template <int N, typename F>
SR_DEVINL void ll_wait(uint4 (&v)[N], F addr, uint32_t gen) {
  unsigned int spins = 0;
  while (true) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && ll_ok(v[i], gen);
    if (ok) break;
    SR_LL_BACKOFF_SLEEP();
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (!ll_ok(v[i], gen)) v[i] = ll_load(addr(i));
    if (++spins > SR_LL_SPIN_LIMIT) ll_timeout(gen);
  }
}

struct DpRing {
  uint8_t* data;
  uint64_t* full;
  uint64_t* empty;
  int slot;
  uint32_t phase;
  int n;  // stages in use (<= DP_STAGES)
  SR_DEVINL void advance() {
    if (++slot == n) { slot = 0; phase ^= 1; }
  }
};

// ------------------------------------------------------------------------------------------------ producer
// Streams this CTA's units of one packed matrix: unit u = RG consecutive 16-row groups; stage = 16 blocks of 1 KB
// (RG = 1: 16 k-chunks of the group; RG = 2: 8 k-chunks of each of the two groups, second group at +8 KB).
// ONE thread issues every copy of the CTA, so this loop is kept minimal.
// ROT: the stage order of unit u is rotated by (7u mod nst) -- used for the down projection, whose activations are read
// from L2 by every CTA: without the rotation all CTAs request the same lines at the same time (hot L2 slices).
SR_DEVINL int dp_rot(int u, int nst) { return (u * 7) % nst; }
template <int RG, bool ROT>
SR_DEVINL void dp_produce(DpRing& r, int v0, const __nv_bfloat16* Wp, int n_units, int nchunks, uint64_t policy, bool hint, int stride) {
  constexpr int CPS = DP_STAGE_BLOCKS / RG;
  const int nst = (nchunks + CPS - 1) / CPS;
  for (int u = v0; u < n_units; u += stride) {
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
      if (ROT && ++st == nst) st = 0;
      if (!ROT) ++st;
    }
  }
}

// Fused gate/up + down schedule: for every gate/up unit of this CTA (32 packed rows = 16 SwiGLU columns, 8 ring stages of two
// 8 KB halves) directly followed by the K-slice of the down projection that consumes exactly those 16 columns (2048 x 16 =
// 64 KB contiguous = 4 ring stages).
SR_DEVINL void dp_produce_de(DpRing& r, int v0, const __nv_bfloat16* Wgu, const __nv_bfloat16* Wdk, int n_units, int nchunks, uint64_t policy,
                             bool hint, int stride) {
  constexpr int CPS = DP_STAGE_BLOCKS / 2;
  const int nst = (nchunks + CPS - 1) / CPS;
  auto issue = [&](uint8_t* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if (hint)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                       smem_u32(dst)),
                   "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                   : "memory");
    else
      dp_bulk_g2s(dst, src, bytes, bar);
  };
  for (int u = v0; u < n_units; u += stride) {
    for (int st = 0; st < nst; ++st) {
      mbar_wait(&r.empty[r.slot], r.phase ^ 1);
      const int nch = min(CPS, nchunks - st * CPS);
      mbar_arrive_expect_tx(&r.full[r.slot], static_cast<uint32_t>(2 * nch * 1024));
#pragma unroll
      for (int g = 0; g < 2; ++g)
        issue(r.data + r.slot * DP_STAGE_BYTES + g * (CPS * 1024), Wgu + (static_cast<int64_t>(u * 2 + g) * nchunks + st * CPS) * 512,
              static_cast<uint32_t>(nch * 1024), &r.full[r.slot]);
      r.advance();
    }
    for (int st = 0; st < 4; ++st) {
      mbar_wait(&r.empty[r.slot], r.phase ^ 1);
      mbar_arrive_expect_tx(&r.full[r.slot], static_cast<uint32_t>(DP_STAGE_BYTES));
      issue(r.data + r.slot * DP_STAGE_BYTES, Wdk + (static_cast<int64_t>(u) * 4 + st) * (DP_STAGE_BYTES / 2), DP_STAGE_BYTES, &r.full[r.slot]);
      r.advance();
    }
  }
}

// ------------------------------------------------------------------------------------------------ consumers
// Stage M (<= 8) rows of width DP_H from an LL buffer into shared memory, warp cw <-> row cw, optional RMSNorm
// (y = w * x * rsqrt(mean x^2 + eps), one rounding to bf16 -- the arithmetic of rmsnorm_kernel in sr_elementwise.cu).
// All 16 lines of a lane are requested at once; only lines that are not there yet are polled again.  Rows >= M are zero.
template <bool NORM>
SR_DEVINL void dp_stage_rows(uint8_t* xs, const uint4* src, uint32_t gen, int M, const float* nw, float eps, int cw, int lane) {
  constexpr int NL = DP_HL / 32;  // lines per lane
  uint4 v[NL];
  float4 w4[NORM ? NL : 1];
  const bool ok = cw < M;
  const uint4* row = src + cw * DP_HL + lane;
  if (NORM) {  // the norm weights do not depend on x: requested together with it
#pragma unroll
    for (int i = 0; i < NL; ++i) w4[i] = __ldg(reinterpret_cast<const float4*>(nw + (lane + 32 * i) * 4));
  }
  if (ok) {
#pragma unroll
    for (int i = 0; i < NL; ++i) v[i] = ll_load(row + 32 * i);
    ll_wait(v, [&](int i) { return row + 32 * i; }, gen);
  } else {
#pragma unroll
    for (int i = 0; i < NL; ++i) v[i] = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  if (NORM) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const float a = bf16_lo(v[i].x), b = bf16_hi(v[i].x), c = bf16_lo(v[i].z), d = bf16_hi(v[i].z);
      ss += a * a + b * b + c * c + d * d;
    }
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss / static_cast<float>(DP_H) + eps);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const float4 w = w4[i];
      v[i].x = pack_bf16x2(bf16_lo(v[i].x) * inv * w.x, bf16_hi(v[i].x) * inv * w.y);
      v[i].z = pack_bf16x2(bf16_lo(v[i].z) * inv * w.z, bf16_hi(v[i].z) * inv * w.w);
    }
  }
#pragma unroll
  for (int i = 0; i < NL; ++i) *reinterpret_cast<uint2*>(xs + cw * DP_XS_STRIDE + (lane + 32 * i) * 8) = make_uint2(v[i].x, v[i].z);
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

enum { DP_EPI_BIAS = 0, DP_EPI_RESIDUAL = 1, DP_EPI_SWIGLU = 2, DP_EPI_F32 = 3 };

struct DpOut {
  uint4* ll;             // LL destination (row-major, `row_lines` lines per batch row); nullptr for the fp32 epilogue
  int row_lines;
  uint32_t gen;
  const uint4* res_ll;   // residual source (DP_HL lines per row), DP_EPI_RESIDUAL only
  uint32_t res_gen;
  const float* bias;
  float* f32;            // fp32 destination (LM head)
  int64_t ld_f32;
  __nv_bfloat16* dbg;    // optional plain bf16 dump
  int64_t ld_dbg;
};

// One GEMV phase: all units of this CTA.  STAGED: activations come from xs (K = DP_H), else plain bf16 from xg (L2).
// Epilogue: thread (row group q, batch b, feature quad fg) sums the 8 warps' partials of 4 consecutive output columns and writes
// one LL line (4 bf16) -- or 4 fp32 logits.
template <int RG, int EPI, bool STAGED>
SR_DEVINL void dp_gemv_phase(DpRing& r, int v0, int n_units, int nchunks, const uint8_t* xs, const __nv_bfloat16* xg, int64_t ldxg,
                             int M, float* red, int& red_par, const DpOut& o, int cw, int lane, int stride) {
  const int g = lane >> 2, t = lane & 3;
  const int ct = threadIdx.x - 32;
  constexpr int NG = (EPI == DP_EPI_SWIGLU) ? 1 : RG;   // independent 16-column groups per unit
  const int eq = ct >> 5, eb = ct & 7, efg = (ct >> 3) & 3;
  const bool epi_thread = ct < 32 * NG && eb < M;
  for (int u = v0; u < n_units; u += stride) {
    float acc[RG][4];
#pragma unroll
    for (int q = 0; q < RG; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
    // epilogue operands of this thread's output columns: requested before the K loop so their latency is hidden
    const int col = (u * NG + eq) * 16 + efg * 4;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 resl = make_uint4(0, 0, 0, 0);
    if (epi_thread) {
      if (EPI == DP_EPI_BIAS && o.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(o.bias + col));
      if (EPI == DP_EPI_RESIDUAL) resl = ll_load(o.res_ll + eb * DP_HL + (col >> 2));
    }
    static_assert(STAGED, "activations are always staged in shared memory");
    dp_consume_staged<RG>(r, nchunks, xs, acc, cw, lane);
    // cross-warp reduction: C fragment (feature g | g+8, batch 2t | 2t+1) -> red[par][warp][q][batch * 20 + feature]
    float* rp = red + red_par * (DP_CONSUMERS * 2 * DP_RED_WQ);
#pragma unroll
    for (int q = 0; q < RG; ++q) {
      float* d = rp + (cw * 2 + q) * DP_RED_WQ;
      d[(2 * t) * 20 + g] = acc[q][0];
      d[(2 * t + 1) * 20 + g] = acc[q][1];
      d[(2 * t) * 20 + g + 8] = acc[q][2];
      d[(2 * t + 1) * 20 + g + 8] = acc[q][3];
    }
    dp_cbar();
    if (epi_thread) {
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
      for (int w8 = 0; w8 < DP_CONSUMERS; ++w8) {
        const float4 a = *reinterpret_cast<const float4*>(rp + (w8 * 2 + (EPI == DP_EPI_SWIGLU ? 0 : eq)) * DP_RED_WQ + eb * 20 + efg * 4);
        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        if (EPI == DP_EPI_SWIGLU) {
          const float4 b = *reinterpret_cast<const float4*>(rp + (w8 * 2 + 1) * DP_RED_WQ + eb * 20 + efg * 4);
          s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
        }
      }
      float v[4] = {s0.x, s0.y, s0.z, s0.w};
      if (EPI == DP_EPI_SWIGLU) {
        const float up[4] = {s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (v[j] / (1.0f + __expf(-v[j]))) * up[j];
      } else if (EPI == DP_EPI_BIAS) {
        v[0] += bias4.x; v[1] += bias4.y; v[2] += bias4.z; v[3] += bias4.w;
      } else if (EPI == DP_EPI_RESIDUAL) {
        ll_spin(resl, o.res_ll + eb * DP_HL + (col >> 2), o.res_gen);
        v[0] += bf16_lo(resl.x); v[1] += bf16_hi(resl.x); v[2] += bf16_lo(resl.z); v[3] += bf16_hi(resl.z);
      }
      if (EPI == DP_EPI_F32) {
        *reinterpret_cast<float4*>(o.f32 + static_cast<int64_t>(eb) * o.ld_f32 + col) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        const uint32_t d0 = pack_bf16x2(v[0], v[1]), d1 = pack_bf16x2(v[2], v[3]);
        if (o.ll != nullptr) ll_store(o.ll + static_cast<int64_t>(eb) * o.row_lines + (col >> 2), d0, d1, o.gen);
        if (o.dbg != nullptr) *reinterpret_cast<uint2*>(o.dbg + static_cast<int64_t>(eb) * o.ld_dbg + col) = make_uint2(d0, d1);
      }
    }
    red_par ^= 1;  // the next unit writes the other buffer; its barrier orders this unit's reads before the reuse after it
  }
}

// ------------------------------------------------------------------------------------------------ fused gate/up + down phase
// For every gate/up unit of this CTA: [gate|up] = xs . Wgu^T (as dp_gemv_phase<2, SWIGLU>), act = silu(gate) * up for the unit's 16
// columns lands in shared memory (never in global memory), and the down projection's K-slice for exactly these columns is applied
// right away: acc_e[row group][..] += Wd[:, 16u .. 16u+15] . act^T, one mma.m16n8k16 per 16-row group, 16 row groups per warp.  The
// SwiGLU output therefore never crosses CTAs; what crosses is every CTA's fp32 partial of x (2048 x 8), published as LL lines and
// summed in a fixed order by the CTA that owns the 16 output rows (dp_reduce_x).  Compared with "all CTAs read all of act" this
// removes 176 KB of activation reads per CTA and one exchange from the middle of the weight stream.
SR_DEVINL int dp_kpos(int k) { return ((k & 7) >> 1) * 4 + (k >> 3) * 2 + (k & 1); }  // position of k inside a 16-wide k group
SR_DEVINL void dp_de_phase(DpRing& r, int v0, int n_units, int nchunks, const uint8_t* xs, uint8_t* s_act, int M, float* red, int& red_par,
                           __nv_bfloat16* dbg_act, int I, float (&acc_e)[16][4], int cw, int lane, int stride) {
  constexpr int CPS = DP_STAGE_BLOCKS / 2;
  const int g = lane >> 2, t = lane & 3;
  const int ct = threadIdx.x - 32;
  const int eb = ct & 7, efg = (ct >> 3) & 3;
  const bool epi_thread = ct < 32 && eb < M;
  const int nst = (nchunks + CPS - 1) / CPS;
  const uint8_t* xrow = xs + g * DP_XS_STRIDE + t * 16;
  for (int u = v0; u < n_units; u += stride) {
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int st = 0; st < nst; ++st) {
      mbar_wait(&r.full[r.slot], r.phase);
      const uint8_t* stage = r.data + r.slot * DP_STAGE_BYTES;
      const int c = st * CPS + cw;
      if (c < nchunks) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xrow + c * 64);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint8_t* blk = stage + q * (CPS * 1024) + cw * 1024;
          const uint4 w0 = *reinterpret_cast<const uint4*>(blk + g * 64 + t * 16);
          const uint4 w1 = *reinterpret_cast<const uint4*>(blk + (g + 8) * 64 + t * 16);
          dp_mma(acc[q], w0.x, w1.x, w0.y, w1.y, xv.x, xv.y);
          dp_mma(acc[q], w0.z, w1.z, w0.w, w1.w, xv.z, xv.w);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&r.empty[r.slot]);
      r.advance();
    }
    float* rp = red + red_par * (DP_CONSUMERS * 2 * DP_RED_WQ);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float* d = rp + (cw * 2 + q) * DP_RED_WQ;
      d[(2 * t) * 20 + g] = acc[q][0];
      d[(2 * t + 1) * 20 + g] = acc[q][1];
      d[(2 * t) * 20 + g + 8] = acc[q][2];
      d[(2 * t + 1) * 20 + g + 8] = acc[q][3];
    }
    dp_cbar();
    if (epi_thread) {
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
      for (int w8 = 0; w8 < DP_CONSUMERS; ++w8) {
        const float4 a = *reinterpret_cast<const float4*>(rp + (w8 * 2) * DP_RED_WQ + eb * 20 + efg * 4);
        const float4 b = *reinterpret_cast<const float4*>(rp + (w8 * 2 + 1) * DP_RED_WQ + eb * 20 + efg * 4);
        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
      }
      const float gt[4] = {s0.x, s0.y, s0.z, s0.w}, up[4] = {s1.x, s1.y, s1.z, s1.w};
      __nv_bfloat16 a4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a4[j] = __float2bfloat16((gt[j] / (1.0f + __expf(-gt[j]))) * up[j]);
        reinterpret_cast<__nv_bfloat16*>(s_act)[eb * 16 + dp_kpos(efg * 4 + j)] = a4[j];
      }
      if (dbg_act != nullptr) *reinterpret_cast<uint2*>(dbg_act + static_cast<int64_t>(eb) * I + u * 16 + efg * 4) = *reinterpret_cast<const uint2*>(a4);
    }
    red_par ^= 1;
    dp_cbar();
    // down-projection slice of this unit: B fragment = act (batch g; k quad t), A fragments straight from the ring
    const uint2 bv = *reinterpret_cast<const uint2*>(s_act + g * 32 + t * 8);
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      mbar_wait(&r.full[r.slot], r.phase);
      const uint8_t* wb = r.data + r.slot * DP_STAGE_BYTES + cw * 2048;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 a_lo = *reinterpret_cast<const uint2*>(wb + j * 512 + g * 32 + t * 8);
        const uint2 a_hi = *reinterpret_cast<const uint2*>(wb + j * 512 + (g + 8) * 32 + t * 8);
        dp_mma(acc_e[st * 4 + j], a_lo.x, a_hi.x, a_lo.y, a_hi.y, bv.x, bv.y);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&r.empty[r.slot]);
      r.advance();
    }
  }
}

// Sum of the 148 partials for this CTA's 16 output rows (fixed order: deterministic), + residual, -> next x as LL lines.
SR_DEVINL void dp_reduce_x(const DpParams& p, int ud, uint32_t gen, const uint4* res_ll, uint4* out_ll, float* red4, int ct,
                           __nv_bfloat16* dbg, int64_t ld_dbg) {
  const int G = static_cast<int>(gridDim.x);
  const int item = ct & 63, grp = ct >> 6;            // item = line inside the unit: row = item / 4, batch pair = item % 4
  const int eb = ct & 7, efg = (ct >> 3) & 3;
  uint4 resl = make_uint4(0, 0, 0, 0);
  const uint4* res_ptr = res_ll + eb * DP_HL + ud * 4 + efg;
  if (ct < 32 && eb < p.B) resl = ll_load(res_ptr);
  float s0 = 0.f, s1 = 0.f;
  const uint4* base = p.xpart_ll + static_cast<int64_t>(ud) * 64 + item;
  constexpr int RB = 13;  // partials per thread and round trip (148 CTAs / 4 thread groups = 37 = 3 x 13 - 2)
  for (int p0 = grp; p0 < G; p0 += 4 * RB) {
    uint4 v[RB];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int pi = p0 + 4 * j;
      v[j] = pi < G ? ll_load(base + static_cast<int64_t>(pi) * (DP_H * 4)) : make_uint4(0, gen, 0, gen);
    }
    ll_wait(v, [&](int j) { return base + static_cast<int64_t>(p0 + 4 * j) * (DP_H * 4); }, gen);
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      s0 += __uint_as_float(v[j].x);
      s1 += __uint_as_float(v[j].z);
    }
  }
  red4[(grp * 64 + item) * 2] = s0;
  red4[(grp * 64 + item) * 2 + 1] = s1;
  dp_cbar();
  float* sx = red4 + 4 * 64 * 2;  // [8 batch][16 rows]
  if (ct < 64) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      a += red4[(q * 64 + ct) * 2];
      b += red4[(q * 64 + ct) * 2 + 1];
    }
    const int row = ct >> 2, t = ct & 3;
    sx[(2 * t) * 16 + row] = a;
    sx[(2 * t + 1) * 16 + row] = b;
  }
  dp_cbar();
  if (ct < 32 && eb < p.B) {
    ll_spin(resl, res_ptr, gen);
    const float4 v = *reinterpret_cast<const float4*>(sx + eb * 16 + efg * 4);
    const uint32_t d0 = pack_bf16x2(v.x + bf16_lo(resl.x), v.y + bf16_hi(resl.x)), d1 = pack_bf16x2(v.z + bf16_lo(resl.z), v.w + bf16_hi(resl.z));
    ll_store(out_ll + eb * DP_HL + ud * 4 + efg, d0, d1, gen);
    if (dbg != nullptr) *reinterpret_cast<uint2*>(dbg + static_cast<int64_t>(eb) * ld_dbg + ud * 16 + efg * 4) = make_uint2(d0, d1);
  }
}

// ------------------------------------------------------------------------------------------------ attention phase
// M-RoPE (mcore_adapter/.../modeling_qwen2_5_vl.py:21-41) on q and the new k, KV-cache append, split-KV attention of the
// new token against the cache (the arithmetic of mrope_kv_kernel + decode_attn_kernel in sr_elementwise.cu / sr_tail.cu).
// Unit = (sample, kv head, key split), exactly one per CTA (n_units <= gridDim.x).
//   * kv_start / kv_len / cache_row and the step's cos/sin table sit in shared memory (loaded once per launch);
//   * the K rows and V^T vectors of the OLD keys do not depend on this layer's qkv: they are requested into registers
//     BEFORE q is polled (dp_attn_prefetch) and arrive while the qkv phase of the other CTAs finishes;
//   * q (+ the new k, v of the owning split, which never re-reads them from the cache: they go to shared memory and,
//     fire-and-forget, to the cache) are polled from the qkv LL buffer, then scores / softmax / P.V from registers;
//   * the un-normalised partial (o, m, l) is published as LL lines; split 0 of every (sample, kv head) polls the other splits'
//     partials (all loads in flight at once), merges, and publishes the attention output as LL lines.
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
// Compact formulation on mma.sync (the instruction footprint of a persistent kernel matters: everything beyond the 32 KB
// L1.5 instruction cache is re-fetched from L2 every layer, in competition with 6 TB/s of weight traffic).
// Scores: A = K rows (16 keys per warp, 8 warps = one 128-key pass), B = q (8 heads); P.V: A = V^T rows (16 output dims per
// warp), B = P (bf16, shared memory).  As in the GEMV phases a thread's 16-byte global load IS an A fragment (the k index is
// permuted identically in both operands), so every cache element is requested exactly once, 16 loads per thread and pass.
// Thread (warp w, g = lane / 4, t = lane % 4) owns heads 2t, 2t+1 in both accumulators:
//   scores  c0..c3 = (key 16w+g | 16w+g+8) x (head 2t | 2t+1),   output c0..c3 = (dim 16w+g | 16w+g+8) x (head 2t | 2t+1).
constexpr int DPB_ROW = 320;                           // bytes per bf16 [128] row (+64 B: conflict-free 16 B fragment reads)
constexpr int DPB_Q = 0;                               // bf16 [8 heads][128]: roped q
constexpr int DPB_P = DPB_Q + 8 * DPB_ROW;             // bf16 [8 heads][128 keys]: probabilities of the pass / merged output
constexpr int DPB_RED = DPB_P + 8 * DPB_ROW;           // float [2][8 warps][8 heads]
constexpr int DPB_KNEW = DPB_RED + 2 * 8 * 8 * 4;      // float [128] roped k of the new token (bf16-rounded values)
constexpr int DPB_VNEW = DPB_KNEW + 128 * 4;           // float [128]
constexpr int DPB_SNEW = DPB_VNEW + 128 * 4;           // float [8] score of the new key per head (log2 domain)
constexpr int DPB_MS = DPB_SNEW + 8 * 4;               // merge: float m[32][8], l[32][8], f[32][8]
constexpr int DPB_LL = DPB_MS + 3 * DP_NS_MAX * 8 * 4; // float [8]
constexpr int DPB_END = DPB_LL + 8 * 4;
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

// L2 prefetch of the split's old K rows (256 B each, one line pair per key) and V^T rows (one per dim, keys contiguous): issued at
// the start of the layer, so that the register loads of dp_attn_prefetch (issued right before q is polled, and returning in
// order with the poll loads behind them) come back at L2 instead of HBM latency.
SR_DEVINL void dp_attn_prefetch_l2(const DpParams& p, int layer, const DpAttnUnit& U) {
  const int ke = dp_attn_old_end(U);
  const int n = ke - U.k_begin;
  if (n <= 0) return;
  const int ct = threadIdx.x - 32;
  const __nv_bfloat16* Kc = p.Kc + static_cast<int64_t>(layer) * p.k_layer_stride;
  const __nv_bfloat16* Vt = p.Vt + static_cast<int64_t>(layer) * p.vt_layer_stride;
  for (int k = ct; k < n; k += DP_CTHREADS) {
    const __nv_bfloat16* kr = Kc + (static_cast<int64_t>(U.start + U.k_begin + k) * p.nkv + U.kvh) * 128;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(kr));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(kr + 64));
  }
  if (ct < 128) {
    const __nv_bfloat16* vr = Vt + static_cast<int64_t>(U.kvh * 128 + ct) * p.ldvt + U.start + U.k_begin;
    for (int k = 0; k < n; k += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(vr + k));
  }
}

SR_DEVINL void dp_attention_phase(const DpParams& p, int layer, uint32_t gen, uint8_t* sm, const float* s_cos, const float* s_sin,
                                   const DpAttnUnit& U, DpAttnFrag& f, int tr_base) {
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
  __nv_bfloat16* Kc = p.Kc + static_cast<int64_t>(layer) * p.k_layer_stride;
  __nv_bfloat16* Vt = p.Vt + static_cast<int64_t>(layer) * p.vt_layer_stride;
  const int qrow_lines = (p.nq + 2 * p.nkv) * 32;
  const int b = U.b, kvh = U.kvh;
  const uint4* qrow = p.qkv_ll + static_cast<int64_t>(b) * qrow_lines;
  const float* cs = s_cos + b * 64;
  const float* sn = s_sin + b * 64;
  // one round trip: q of the heads (threads 0..63: head ct/8, dims 8(ct%8).. and +64), new k (64..79), new v (128..159)
  if (ct < 64) {
    const int h = ct >> 3, d0 = (ct & 7) * 8;
    uint32_t lw[4] = {0, 0, 0, 0}, hw[4] = {0, 0, 0, 0};
    if (h < G) {
      const uint4* ql = qrow + (kvh * G + h) * 32 + (d0 >> 2);
      uint4 qv[4] = {ll_load(ql), ll_load(ql + 1), ll_load(ql + 16), ll_load(ql + 17)};
      ll_wait(qv, [&](int i) { return ql + (i & 1) + 16 * (i >> 1); }, gen);
      lw[0] = qv[0].x; lw[1] = qv[0].z; lw[2] = qv[1].x; lw[3] = qv[1].z;
      hw[0] = qv[2].x; hw[1] = qv[2].z; hw[2] = qv[3].x; hw[3] = qv[3].z;
    }
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
  } else if (U.owner && ct < 80) {
    const int d = (ct - 64) * 4;  // dims d..d+3 and 64+d..
    const uint4* kl = qrow + (p.nq + kvh) * 32 + (d >> 2);
    uint4 kv2[2] = {ll_load(kl), ll_load(kl + 16)};
    ll_wait(kv2, [&](int i) { return kl + 16 * i; }, gen);
    const uint4 a = kv2[0], c = kv2[1];
    const float x1[4] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.z), bf16_hi(a.z)};
    const float x2[4] = {bf16_lo(c.x), bf16_hi(c.x), bf16_lo(c.z), bf16_hi(c.z)};
    __nv_bfloat16 k1[4], k2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      k1[j] = __float2bfloat16(x1[j] * cs[d + j] - x2[j] * sn[d + j]);
      k2[j] = __float2bfloat16(x2[j] * cs[d + j] + x1[j] * sn[d + j]);
      s_knew[d + j] = __bfloat162float(k1[j]);
      s_knew[d + j + 64] = __bfloat162float(k2[j]);
    }
    __nv_bfloat16* kd = Kc + (static_cast<int64_t>(U.crow) * p.nkv + kvh) * 128;
    *reinterpret_cast<uint2*>(kd + d) = *reinterpret_cast<const uint2*>(k1);
    *reinterpret_cast<uint2*>(kd + d + 64) = *reinterpret_cast<const uint2*>(k2);
  } else if (U.owner && ct >= 128 && ct < 160) {
    const int d = (ct - 128) * 4;
    const uint4* vl = qrow + (p.nq + p.nkv + kvh) * 32 + (d >> 2);
    uint4 a = ll_load(vl);
    ll_spin(a, vl, gen);
    const uint32_t u2[2] = {a.x, a.z};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned short bits = static_cast<unsigned short>(j & 1 ? u2[j >> 1] >> 16 : u2[j >> 1] & 0xffffu);
      *reinterpret_cast<unsigned short*>(Vt + static_cast<int64_t>(kvh * 128 + d + j) * p.ldvt + U.crow) = bits;
      s_vnew[d + j] = __uint_as_float(static_cast<uint32_t>(bits) << 16);
    }
  }
  dp_cbar();
  if (ct == 0) dp_trace_at(p, 1, tr_base);
  if (p.flags & 8) dp_attn_prefetch(p, layer, U, f);  // diagnostic: K / V^T requested only after q has arrived
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
    if (p0 != U.k_begin) dp_attn_load(p, U, Kc, Vt, p0, nk, f);  // first pass: requested before q was polled
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
  // publish this split's un-normalised partial (o, m, l): line (dim * 4 + t) = o[dim][heads 2t, 2t+1]
  uint4* wo = p.part_ll + (static_cast<int64_t>(U.bk) * ns + U.split) * DP_PART_LINES;
  const int d0 = 16 * w + g;
  ll_store(wo + d0 * 4 + t, __float_as_uint(o[0]), __float_as_uint(o[1]), gen);
  ll_store(wo + (d0 + 8) * 4 + t, __float_as_uint(o[2]), __float_as_uint(o[3]), gen);
  dp_cbar();
  if (ct < 4) {  // lanes t = 0..3 of warp 0 hold the running maxima of heads 2t, 2t+1
    float L0 = 0.f, L1 = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < DP_CONSUMERS; ++w8) {
      L0 += s_red[w8 * 8 + 2 * ct];
      L1 += s_red[w8 * 8 + 2 * ct + 1];
    }
    ll_store(wo + 512 + 2 * ct, __float_as_uint(m0), __float_as_uint(L0), gen);
    ll_store(wo + 512 + 2 * ct + 1, __float_as_uint(m1), __float_as_uint(L1), gen);
  }
  if (ct == 0) dp_trace_at(p, 1, tr_base + 1);
  dp_cbar();  // scratch reads done before the region is reused (merge below / next phase's staging)
  if (U.split != 0) return;
  // ---- merge (split 0 of this (sample, kv head)): statistics and partial outputs polled together
  const uint4* wb = p.part_ll + static_cast<int64_t>(U.bk) * ns * DP_PART_LINES;
  float O[4] = {0.f, 0.f, 0.f, 0.f};
  uint4 st_line = make_uint4(0, 0, 0, 0);
  const bool st_thread = ct < ns * 8;
  const uint4* st_ptr = wb + (ct >> 3) * DP_PART_LINES + 512 + (ct & 7);
  if (st_thread) st_line = ll_load(st_ptr);
  for (int c0 = 0; c0 < ns; c0 += DP_MERGE_CHUNK) {
    uint4 ov[2 * DP_MERGE_CHUNK];  // [split j][dim d0 | d0 + 8]
#pragma unroll
    for (int j = 0; j < DP_MERGE_CHUNK; ++j) {
      if (c0 + j < ns) {
        ov[2 * j] = ll_load(wb + (c0 + j) * DP_PART_LINES + d0 * 4 + t);
        ov[2 * j + 1] = ll_load(wb + (c0 + j) * DP_PART_LINES + (d0 + 8) * 4 + t);
      } else {
        ov[2 * j] = ov[2 * j + 1] = make_uint4(0, gen, 0, gen);
      }
    }
    if (c0 == 0) {
      if (st_thread) {
        ll_spin(st_line, st_ptr, gen);
        m_ms[ct] = __uint_as_float(st_line.x);
        m_ls[ct] = __uint_as_float(st_line.z);
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
    }
    ll_wait(ov, [&](int i) { return wb + (c0 + (i >> 1)) * DP_PART_LINES + (d0 + 8 * (i & 1)) * 4 + t; }, gen);
#pragma unroll
    for (int j = 0; j < DP_MERGE_CHUNK; ++j) {
      if (c0 + j < ns) {
        const float f0 = m_f[(c0 + j) * 8 + 2 * t], f1 = m_f[(c0 + j) * 8 + 2 * t + 1];
        O[0] += __uint_as_float(ov[2 * j].x) * f0;
        O[1] += __uint_as_float(ov[2 * j].z) * f1;
        O[2] += __uint_as_float(ov[2 * j + 1].x) * f0;
        O[3] += __uint_as_float(ov[2 * j + 1].z) * f1;
      }
    }
    __syncwarp();
  }
  {  // normalise -> bf16 [8 heads][128] in shared memory (the P region is free now) -> LL lines
    const float L0 = m_L[2 * t], L1 = m_L[2 * t + 1];
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(sm + DPB_P);
    ob[(2 * t) * (DPB_ROW / 2) + d0] = __float2bfloat16(L0 > 0.f ? O[0] / L0 : 0.f);
    ob[(2 * t + 1) * (DPB_ROW / 2) + d0] = __float2bfloat16(L1 > 0.f ? O[1] / L1 : 0.f);
    ob[(2 * t) * (DPB_ROW / 2) + d0 + 8] = __float2bfloat16(L0 > 0.f ? O[2] / L0 : 0.f);
    ob[(2 * t + 1) * (DPB_ROW / 2) + d0 + 8] = __float2bfloat16(L1 > 0.f ? O[3] / L1 : 0.f);
  }
  dp_cbar();
  {
    const int h = ct >> 5, j = ct & 31;  // head h, dims 4j..4j+3
    if (h < G) {
      const uint2 v = *reinterpret_cast<const uint2*>(sm + DPB_P + h * DPB_ROW + j * 8);
      ll_store(p.ao_ll + static_cast<int64_t>(b) * DP_HL + (kvh * G + h) * 32 + j, v.x, v.y, gen);
      if (p.ao != nullptr) *reinterpret_cast<uint2*>(p.ao + static_cast<int64_t>(b) * DP_H + (kvh * G + h) * 128 + j * 4) = v;
    }
  }
  dp_cbar();
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // per-launch constants -> shared memory (pointer / scalar loads off the critical paths)
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.layers);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_layers);
    for (int i = threadIdx.x; i < p.n_layers * 8; i += DP_THREADS) dst[i] = src[i];
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
    if (threadIdx.x == 8) s_scal[24] = static_cast<int>(*reinterpret_cast<volatile unsigned int*>(p.epoch));
    if (threadIdx.x == 9) s_scal[25] = 0;
    if (threadIdx.x >= 32 && threadIdx.x < 96) reinterpret_cast<uint32_t*>(smem + DP_OFF_ACT)[threadIdx.x - 32] = 0u;
  }
  if (threadIdx.x == 0) {
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
  const int n_head = p.vocab / 16;
  const int limit = p.stop_after > 0 ? p.stop_after : 0x7fffffff;
  // Unit -> CTA mapping, rotated per matrix: unit u of a matrix goes to CTA (u + offset) mod gridDim.  The unit counts (80 / 128
  // / 688 / 128 on 148 CTAs) leave some CTAs without a unit, or with one unit more; the offsets make those sets complement each
  // other (e.g. the 20 CTAs idle in the down projection take a fifth gate/up unit), which lowers the bytes of the busiest CTA.
  auto vidx = [&](int off) { const int g = static_cast<int>(gridDim.x); return (static_cast<int>(blockIdx.x) + g - (off % g)) % g; };
  const int vq = vidx(p.unit_off[0]), vo = vidx(p.unit_off[1]), vg = vidx(p.unit_off[2]), vd = vidx(p.unit_off[3]);
  const int G = static_cast<int>(gridDim.x);
  // measurement aid (flags bit 5): the LM head is streamed by every second CTA only -- does one SM ingest faster when fewer stream?
  const bool half_head = (p.flags & 32) != 0;
  const int hs = half_head ? (G + 1) / 2 : G;
  const int hv = half_head ? ((blockIdx.x & 1) ? 0x7fffffff : static_cast<int>(blockIdx.x >> 1)) : static_cast<int>(blockIdx.x);

  if (warp == 0) {
    // ---------------- producer: weights only; runs ahead of the consumers across phases ----------------
    if (lane == 0) {
      uint64_t policy = 0;
      const bool hint = (p.flags & 1) != 0;
      if (hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      int ph = 0;
      dp_trace_at(p, 0, 0);
      for (int l = 0; l < p.n_layers; ++l) {
        const DpLayer L = s_layers[l];
        if (ph++ >= limit) return;
        dp_produce<2, false>(ring, vq, L.qkv_wp, n_qkv / 2, kq, policy, hint, G);
        dp_trace_at(p, 0, 1 + 4 * l);
        if (ph++ >= limit) return;  // attention phase: no weights
        if (ph++ >= limit) return;
        dp_produce<1, false>(ring, vo, L.o_wp, n_o, kq, policy, hint, G);
        dp_trace_at(p, 0, 2 + 4 * l);
        if (ph++ >= limit) return;
        if (p.flags & 4) {  // diagnostic: no gate/up streaming before this CTA's consumers reach phase D (idle HBM during the chain)
          volatile int* gate = s_scal + 25;
          while (*gate < l + 1) __nanosleep(20);
        }
        dp_produce_de(ring, vg, L.gu_wp, L.down_wk, n_gu, kq, policy, hint, G);
        dp_trace_at(p, 0, 3 + 4 * l);
        if (ph++ >= limit) return;
        dp_trace_at(p, 0, 4 + 4 * l);
      }
      if (ph++ >= limit) return;
      if (hv < n_head) dp_produce<1, false>(ring, hv, p.lm_head_p, n_head, kq, policy, hint, hs);
      dp_trace_at(p, 0, 1 + 4 * p.n_layers);
    }
    return;
  }
  // ---------------- consumers ----------------
  const int cw = warp - 1;
  const int ct = threadIdx.x - 32;
  const uint32_t epoch_raw = static_cast<uint32_t>(s_scal[24]);
  const uint32_t epoch = epoch_raw + 1u;  // generations epoch .. epoch + n_layers: never 0 on a zero-initialised workspace
  const bool dump = (p.flags & 2) != 0;
  {  // the step's input rows (embedded tokens, plain bf16) -> LL lines of generation `epoch + 0`, spread over all CTAs
    const int line = static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) * ct;
    if (line < p.B * DP_HL) {
      const int b = line / DP_HL, j = line % DP_HL;
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(p.x + static_cast<int64_t>(b) * p.ldx + j * 4));
      ll_store(p.xb_ll + line, v.x, v.y, epoch);
    }
  }
  int red_par = 0;
  int ph = 0;
  const bool tracer = ct == 0;
  const DpAttnUnit U = dp_attn_unit(p, s_scal);
  DpAttnFrag fr;
  constexpr int EV = 9;  // trace events per layer (after event 0)
#define DP_TRACE(k) do { if (tracer) dp_trace_at(p, 1, 1 + EV * l + (k)); } while (0)
  if (tracer) dp_trace_at(p, 1, 0);
  int l = 0;
  for (; l < p.n_layers; ++l) {
    const DpLayer L = s_layers[l];
    const uint32_t gin = epoch + static_cast<uint32_t>(l);      // generation of this layer's input x (written by layer l-1)
    const uint32_t gen = gin + 1;                               // generation of everything this layer writes
    DpOut o;
    if ((p.flags & 16) && U.active) dp_attn_prefetch_l2(p, l, U);  // this layer's K / V^T slice -> L2 while the qkv phase runs
    // A: qkv = rmsnorm(x; ln1) . Wqkv^T + b   (units of 32 rows: 80 double units, one pass)
    if (ph++ >= limit) break;
    if (vq < n_qkv / 2) {
      dp_stage_rows<true>(xs, p.xb_ll, gin, p.B, L.ln1, p.eps, cw, lane);
      dp_cbar();
      DP_TRACE(0);
      o = DpOut{p.qkv_ll, n_qkv * 4, gen, nullptr, 0, L.qkv_b, nullptr, 0, p.qkv, static_cast<int64_t>(n_qkv) * 16};
      dp_gemv_phase<2, DP_EPI_BIAS, true>(ring, vq, n_qkv / 2, kq, xs, nullptr, 0, p.B, red, red_par, o, cw, lane, G);
    } else {
      DP_TRACE(0);
    }
    DP_TRACE(1);
    // B: rope + KV append + attention -> ao
    if (ph++ >= limit) break;
    if (U.active) {
      if (!(p.flags & 8)) dp_attn_prefetch(p, l, U, fr);  // old K / V^T rows fly while q is being produced
      dp_cbar();                      // (phase A's shared-memory reads are done before the scratch region is written)
      dp_attention_phase(p, l, gen, xs, s_cos, s_sin, U, fr, 1 + EV * l + 2);
    } else {
      DP_TRACE(2);
      DP_TRACE(3);
    }
    // C: x += ao . Wo^T
    if (ph++ >= limit) break;
    if (vo < n_o) {
      dp_stage_rows<false>(xs, p.ao_ll, gen, p.B, nullptr, 0.f, cw, lane);
      dp_cbar();
      DP_TRACE(4);
      o = DpOut{p.xa_ll, DP_HL, gen, p.xb_ll, gin, nullptr, nullptr, 0, dump ? p.x : nullptr, p.ldx};
      dp_gemv_phase<1, DP_EPI_RESIDUAL, true>(ring, vo, n_o, kq, xs, nullptr, 0, p.B, red, red_par, o, cw, lane, G);
    } else {
      DP_TRACE(4);
    }
    DP_TRACE(5);
    // D+E: act = silu(gate) * up of this CTA's gate/up units, immediately multiplied by the matching K-slices of Wdown
    if (ph++ >= limit) break;
    dp_cbar();
    dp_stage_rows<true>(xs, p.xa_ll, gen, p.B, L.ln2, p.eps, cw, lane);
    dp_cbar();
    DP_TRACE(6);
    if (tracer) *reinterpret_cast<volatile int*>(s_scal + 25) = l + 1;
    {
      float acc_e[16][4];
#pragma unroll
      for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc_e[i][j] = 0.f;
      dp_de_phase(ring, vg, n_gu, kq, xs, smem + DP_OFF_ACT, p.B, red, red_par, dump ? p.act : nullptr, p.I, acc_e, cw, lane, G);
      // publish this CTA's partial of Wdown . act^T: line (row, batch pair t) = (x[2t][row], x[2t+1][row])
      uint4* part = p.xpart_ll + static_cast<int64_t>(blockIdx.x) * (DP_H * 4);
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = (16 * cw + i) * 16 + g;
        ll_store(part + row * 4 + t, __float_as_uint(acc_e[i][0]), __float_as_uint(acc_e[i][1]), gen);
        ll_store(part + (row + 8) * 4 + t, __float_as_uint(acc_e[i][2]), __float_as_uint(acc_e[i][3]), gen);
      }
    }
    DP_TRACE(7);
    // E': x += sum over CTAs of the partials (rows of this CTA's unit), fixed order
    if (ph++ >= limit) break;
    if (vd < n_o) dp_reduce_x(p, vd, gen, p.xa_ll, p.xb_ll, reinterpret_cast<float*>(smem + DP_OFF_RED4), ct, dump ? p.x : nullptr, p.ldx);
    DP_TRACE(8);
    dp_cbar();  // phase E's epilogue reads of `red` / phase D's of xs are done before the next layer restages
  }
  // final norm + LM head -> fp32 logits
  if (l == p.n_layers && ph++ < limit) {
    dp_stage_rows<true>(xs, p.xb_ll, epoch + static_cast<uint32_t>(p.n_layers), p.B, p.final_norm, p.eps, cw, lane);
    dp_cbar();
    DpOut o = DpOut{nullptr, 0, 0, nullptr, 0, nullptr, p.logits, p.ld_logits, nullptr, 0};
    if (hv < n_head) dp_gemv_phase<1, DP_EPI_F32, true>(ring, hv, n_head, kq, xs, nullptr, 0, p.B, red, red_par, o, cw, lane, hs);
    DP_TRACE(0);
  }
#undef DP_TRACE
  // the last CTA to leave advances the epoch (stream order makes it visible to the next launch / graph replay)
  if (ct == 0) {
    const unsigned int old = atomicAdd(p.epoch + 1, 1u);
    if (old == gridDim.x - 1) {
      p.epoch[1] = 0;
      p.epoch[0] = epoch_raw + static_cast<uint32_t>(p.n_layers) + 2u;
    }
  }
}

}  // namespace sr

using namespace sr;

static inline int64_t dp_align(int64_t v) { return (v + 255) / 256 * 256; }
struct DpWsLayout {
  int64_t xa, xb, qkv, ao, part, xpart, total;
};
static DpWsLayout dp_ws_layout(int n_q_heads, int n_kv_heads, int intermediate) {
  DpWsLayout w;
  int64_t off = 256;  // header: epoch, exit counter
  w.xa = off; off += dp_align(8LL * DP_HL * 16);
  w.xb = off; off += dp_align(8LL * DP_HL * 16);
  w.qkv = off; off += dp_align(8LL * (n_q_heads + 2 * n_kv_heads) * 32 * 16);
  w.ao = off; off += dp_align(8LL * DP_HL * 16);
  (void)intermediate;
  w.part = off; off += dp_align(8LL * n_kv_heads * DP_NS_MAX * DP_PART_LINES * 16);
  w.xpart = off; off += dp_align(256LL * DP_H * 4 * 16);   // up to 256 CTAs x (2048 rows x 4 lines)
  w.total = off;
  return w;
}

extern "C" int64_t sr_decode_persistent_workspace_bytes(int n_q_heads, int n_kv_heads, int intermediate) {
  return dp_ws_layout(n_q_heads, n_kv_heads, intermediate).total;
}

static unsigned long long* g_dp_trace = nullptr;
static int g_dp_trace_events = 0;
extern "C" int sr_set_decode_persistent_trace(void* buffer, int events_per_role) {
  g_dp_trace = reinterpret_cast<unsigned long long*>(buffer);
  g_dp_trace_events = buffer ? events_per_role : 0;
  return SR_OK;
}

static int g_dp_flags = -1;      // -1: env SR_DECODE_FLAGS
static int g_dp_ring = 0;        // 0: env SR_DECODE_RING (default below)
static int g_dp_unit_off[4] = {-1, 0, 0, 0};  // -1: env SR_DECODE_UNIT_OFFSETS="q,o,g,d" (default below)
extern "C" int sr_set_decode_persistent_tuning(int flags, int ring_stages, int off_qkv, int off_o, int off_gate_up, int off_down) {
  if (ring_stages == 1 || ring_stages > DP_STAGES || ring_stages < 0) return SR_ERR_INVALID;
  if (flags >= 0) g_dp_flags = flags & 0xff;
  g_dp_ring = ring_stages;
  if (off_qkv >= 0 && off_o >= 0 && off_gate_up >= 0 && off_down >= 0) {
    g_dp_unit_off[0] = off_qkv;
    g_dp_unit_off[1] = off_o;
    g_dp_unit_off[2] = off_gate_up;
    g_dp_unit_off[3] = off_down;
  }
  return SR_OK;
}

extern "C" int sr_decode_step_persistent(const void* layer_table, int n_layers, const void* lm_head_packed, const float* final_norm,
                                         int vocab, void* x, int64_t ldx, void* qkv_scratch, void* ao_scratch, void* act_scratch,
                                         float* logits, int64_t ld_logits, const float* cosb, const float* sinb,
                                         const int32_t* cache_row, const int32_t* kv_start, const int32_t* kv_len, void* Kc,
                                         int64_t k_layer_stride, void* Vt, int64_t vt_layer_stride, int64_t ldvt, int batch,
                                         int n_q_heads, int n_kv_heads, int intermediate, float norm_eps, float scale,
                                         void* workspace, int stop_after_phases, void* stream) {
  SR_REQUIRE(layer_table && lm_head_packed && final_norm && x && logits && cosb && sinb && cache_row && kv_start && kv_len && Kc && Vt &&
                 workspace,
             "sr_decode_step_persistent: null pointer");
  SR_REQUIRE(batch > 0 && batch <= 8, "sr_decode_step_persistent: batch must be in [1, 8] (got %d)", batch);
  SR_REQUIRE(n_q_heads * 128 == DP_H && n_kv_heads > 0 && n_q_heads % n_kv_heads == 0 && n_q_heads / n_kv_heads <= DP_GMAX,
             "sr_decode_step_persistent: needs hidden 2048 = n_q_heads * 128 and <= 8 query heads per kv head");
  SR_REQUIRE(n_layers <= DP_MAX_LAYERS, "sr_decode_step_persistent: at most %d layers", DP_MAX_LAYERS);
  SR_REQUIRE(intermediate % 32 == 0 && vocab % 16 == 0 && n_layers > 0, "sr_decode_step_persistent: intermediate %% 32, vocab %% 16");
  SR_REQUIRE(ldx % 8 == 0 && ldvt % 8 == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
                 (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(lm_head_packed) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && ld_logits % 4 == 0,
             "sr_decode_step_persistent: alignment");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static int grid_by_dev[64] = {0};  // per device: the function attribute and the SM count belong to the current device
  int dev = 0;
  SR_CHECK_CUDA(cudaGetDevice(&dev));
  SR_REQUIRE(dev >= 0 && dev < 64, "sr_decode_step_persistent: device ordinal %d", dev);
  if (grid_by_dev[dev] == 0) {
    SR_CHECK_CUDA(cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_BYTES));
    int per_sm = 0;
    SR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_persistent_kernel, DP_THREADS, DP_SMEM_BYTES));
    SR_REQUIRE(per_sm >= 1, "sr_decode_step_persistent: kernel does not fit on an SM (smem %d B)", DP_SMEM_BYTES);
    const int n_sm = sm_count();  // one CTA per SM: all co-resident (the CTAs wait for each other's data)
    SR_REQUIRE(n_sm <= 1024 && n_sm <= DP_CTHREADS, "sr_decode_step_persistent: at most %d SMs", DP_CTHREADS);
    grid_by_dev[dev] = n_sm;
  }
  const int grid_cached = grid_by_dev[dev];
  DpParams p;
  p.layers = reinterpret_cast<const DpLayer*>(layer_table);
  p.n_layers = n_layers;
  p.lm_head_p = reinterpret_cast<const __nv_bfloat16*>(lm_head_packed);
  p.final_norm = final_norm;
  p.vocab = vocab;
  p.x = reinterpret_cast<__nv_bfloat16*>(x);
  p.ldx = ldx;
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
  const DpWsLayout w = dp_ws_layout(n_q_heads, n_kv_heads, intermediate);
  char* ws = reinterpret_cast<char*>(workspace);
  p.epoch = reinterpret_cast<unsigned int*>(ws);
  p.xa_ll = reinterpret_cast<uint4*>(ws + w.xa);
  p.xb_ll = reinterpret_cast<uint4*>(ws + w.xb);
  p.qkv_ll = reinterpret_cast<uint4*>(ws + w.qkv);
  p.ao_ll = reinterpret_cast<uint4*>(ws + w.ao);
  p.part_ll = reinterpret_cast<uint4*>(ws + w.part);
  p.xpart_ll = reinterpret_cast<uint4*>(ws + w.xpart);
  SR_REQUIRE(bk <= grid_cached, "sr_decode_step_persistent: batch * n_kv_heads must not exceed the SM count");
  int ns = grid_cached / static_cast<int>(bk);
  p.ns = ns < 1 ? 1 : (ns > DP_NS_MAX ? DP_NS_MAX : ns);
  p.stop_after = stop_after_phases;
  if (g_dp_flags < 0) {
    const char* e = getenv("SR_DECODE_FLAGS");
    g_dp_flags = e ? atoi(e) : 17;  // evict_first on the weight stream (keeps activations and instructions in L2) + K/V L2 prefetch
  }
  static int ring_stages = -1;
  if (ring_stages < 0) {
    const char* e = getenv("SR_DECODE_RING");
    ring_stages = e ? atoi(e) : 8;
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
  p.flags = g_dp_flags;
  const bool dump = (p.flags & 2) != 0;
  SR_REQUIRE(!dump || (qkv_scratch && ao_scratch && act_scratch), "sr_decode_step_persistent: dump flag needs the three scratch buffers");
  p.qkv = dump ? reinterpret_cast<__nv_bfloat16*>(qkv_scratch) : nullptr;
  p.ao = dump ? reinterpret_cast<__nv_bfloat16*>(ao_scratch) : nullptr;
  p.act = dump ? reinterpret_cast<__nv_bfloat16*>(act_scratch) : nullptr;
  p.trace = g_dp_trace;
  p.trace_events = g_dp_trace_events;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_cached);
  cfg.blockDim = dim3(DP_THREADS);
  cfg.dynamicSmemBytes = DP_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // guarantees co-residency of the 148 CTAs (they poll each other's lines)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_persistent_kernel, p));
  return SR_OK;
}
