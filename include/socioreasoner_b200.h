/* socioreasoner_b200.h -- C ABI of the B200-native SocioReasoner hot path (libsocioreasoner_b200.so).
 *
 * The reference (AMAP-ML/SocioReasoner) has NO native/FFI interface: its seam is the Python strategy
 * class hierarchy (roll/distributed/strategy/strategy.py:16-138) and every op below is reached there
 * through a pip library (torch/cuBLAS, flash-attn, HF transformers eager code).  Each entry point cites
 * the reference call site whose arithmetic it replaces; `HF:` = transformers/models/qwen2_5_vl/
 * modeling_qwen2_5_vl.py (the third-party module the reference imports at
 * mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py:190-200 and
 * roll/models/model_providers.py:123-127).
 *
 * Conventions: plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless a name ends
 * in `_host`.  `stream` is a cudaStream_t passed as void*.  Return 0 on success, negative on error
 * (sr_last_error() returns a thread-local message).  No hidden allocations, no global state besides
 * per-process caches of driver entry points / function attributes.  bf16 = __nv_bfloat16 storage,
 * fp32 accumulation.  Every kernel is sm_100a-only; there is no CPU or non-Blackwell fallback.
 */
#ifndef SOCIOREASONER_B200_H
#define SOCIOREASONER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR_ABI_VERSION 1

/* error codes */
#define SR_OK 0
#define SR_ERR_INVALID (-1)
#define SR_ERR_CUDA (-2)
#define SR_ERR_DRIVER (-3)
#define SR_ERR_UNSUPPORTED (-4)

/* GEMM epilogues */
#define SR_EPI_NONE 0     /* C = A.W^T (+bias)                                            */
#define SR_EPI_RESIDUAL 1 /* C = A.W^T (+bias) + residual                                 */
#define SR_EPI_SWIGLU 2   /* C[:, j] = silu(gate_j) * up_j ; W rows packed [16 gate|16 up] */
#define SR_EPI_GELU 3     /* C = gelu_erf(A.W^T + bias)                                   */

#define SR_DTYPE_BF16 0
#define SR_DTYPE_F32 1

/* ---- library ------------------------------------------------------------------------------- */
const char* sr_last_error(void);
int sr_abi_version(void);
/* 0 if the current CUDA device is compute capability 10.x (B200), else SR_ERR_UNSUPPORTED. */
int sr_device_check(void);
/* Programmatic dependent launch for the decode-step kernels (default on; env SR_PDL=0 disables). */
int sr_set_pdl(int enabled);

/* ---- dense contraction (tcgen05 + TMA) ------------------------------------------------------
 * C[M, n_out] = epilogue(A[M,K] . W[N,K]^T), bf16 inputs, fp32 accumulate in TMEM.
 * Replaces torch.nn.functional.linear / Conv3d-as-GEMM at HF:91-114 (patch embed), HF:214,285 (ViT qkv,
 * proj), HF:77-88 (ViT MLP), HF:133-146 (merger), HF:704-707 (decoder q/k/v/o), HF:611-624 (decoder MLP),
 * HF:1385 (lm_head).  lda/ldw/ldc/ldr are row strides in ELEMENTS (multiples of 8).  bias: fp32[N] or NULL.
 * SR_EPI_SWIGLU: W/bias rows are packed in groups of 32 = [16 gate rows | 16 up rows]; n_out = N/2.
 * block_n: 0 = auto (M <= 32 -> sr_gemm_skinny_bf16; 256 for M > 64, else 32), or 32 / 128 / 256 to force a tcgen05 tile.
 * Output / residual rows whose start and row stride are 32-byte aligned (ldc % 16 == 0 for bf16, N % 16 == 0) are written /
 * read with 256-bit accesses; anything else takes 16-byte accesses (same results). */
int sr_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, const float* bias,
                 const void* residual, int64_t ldr, int M, int N, int K, int epilogue, int out_dtype, int block_n,
                 void* stream);

/* Skinny variant for the decode step (M <= 32 live sequences, K % 32 == 0): HBM-bound weight streaming
 * (16-byte ld.global.nc vectors -> mma.sync m16n8k16 fragments, in-CTA split-K), same epilogues.  The decode-time
 * nn.Linear calls of HF:704-707,611-624,1385 that the reference runs inside vLLM (vllm_strategy.py:114-141).
 * sr_gemm_bf16 dispatches here automatically when block_n == 0 and the shape qualifies. */
int sr_gemm_skinny_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, const float* bias,
                        const void* residual, int64_t ldr, int M, int N, int K, int epilogue, int out_dtype, void* stream);

/* Same with RMSNorm fused into the operand load: C = epilogue(rmsnorm(A; norm_weight, eps) . W^T).  Replaces the
 * input_layernorm / post_attention_layernorm / final norm + following nn.Linear pairs of the decode step
 * (HF:57-71 + HF:704-707 / 611-624 / 1385) with one launch. */
int sr_gemm_skinny_norm_bf16(const void* A, int64_t lda, const float* norm_weight, float norm_eps, const void* W, int64_t ldw,
                             void* C, int64_t ldc, const float* bias, const void* residual, int64_t ldr, int M, int N, int K,
                             int epilogue, int out_dtype, void* stream);

/* Same on weights pre-packed for streaming: Wpacked[N/16][K/32][16][32] (block (g,c) holds rows 16g..16g+15, columns
 * 32c..32c+31, 1 KB contiguous; N % 16 == 0, K % 32 == 0).  A warp-wide 16-byte load then covers 512 contiguous bytes and a
 * CTA streams one linear region of HBM.  norm_weight may be NULL (no fused RMSNorm). */
int sr_gemm_skinny_packed_bf16(const void* A, int64_t lda, const float* norm_weight, float norm_eps, const void* Wpacked, void* C,
                               int64_t ldc, const float* bias, const void* residual, int64_t ldr, int M, int N, int K,
                               int epilogue, int out_dtype, void* stream);

/* Variant of sr_gemm_skinny_packed_bf16 for M <= 8 that streams the packed weights with the bulk-copy engine
 * (cp.async.bulk into a 96-128 KB shared-memory ring, producer/consumer mbarriers).  Measured on B200 it is NOT faster than
 * the register-streaming kernel (tools/gemv_bench.py: 2.5-4.1 vs 3.0-4.7 TB/s), so sr_gemm_skinny_packed_bf16 uses it only
 * after sr_set_gemv_bulk(1) / env SR_GEMV_BULK=1. */
int sr_gemm_skinny_bulk_bf16(const void* A, int64_t lda, const float* norm_weight, float norm_eps, const void* Wpacked, void* C,
                             int64_t ldc, const float* bias, const void* residual, int64_t ldr, int M, int N, int K,
                             int epilogue, int out_dtype, void* stream);
int sr_set_gemv_bulk(int enabled);
/* Split-K variant for long-K / narrow-N decode linears (down_proj: N=2048 gives only 128 CTAs): two CTAs per 16-row group
 * each stream half of K, fp32 partials meet in `workspace` (sr_gemm_skinny_splitk_workspace_bytes(N) bytes, the leading
 * ceil(N/16) int32 tickets zeroed once by the caller; the kernel re-arms them).  M <= 8, no fused norm, no SwiGLU. */
int64_t sr_gemm_skinny_splitk_workspace_bytes(int N);
int sr_gemm_skinny_packed_splitk_bf16(const void* A, int64_t lda, const void* Wpacked, void* C, int64_t ldc, const float* bias,
                                      const void* residual, int64_t ldr, int M, int N, int K, int epilogue, int out_dtype,
                                      void* workspace, void* stream);
/* Tuning knob: warps per CTA of the skinny GEMM at M <= 8 (0 = default, 4 / 8 / 16). */
int sr_set_gemv_warps(int warps);
/* Tuning knob: weight-load instruction flavour of the skinny GEMM at M <= 8 (0 nc/no-allocate, 1 + L2::256B, 2 plain). */
int sr_set_gemv_load(int flavour);

/* Attention kernel variant (env SR_ATTN_LEAN; default 2):
 *   2 = two CTAs per SM, P and O in tensor memory (TS-form P.V MMA, lazy softmax reference point), K reloaded during the
 *       softmax, double-buffered V, persistent grid for plans with many equally long work items;
 *   3 = variant 2 with one CTA per work item (A/B measurements; same bits as 2);
 *   1 = round-1 lean kernel: two CTAs per SM, single K/V stage, P in shared memory aliasing K;
 *   0 = round-1 kernel with one CTA per SM, 2-stage K/V ring and double-buffered S (same bits as 1).
 * Variants 2 / 3 agree with 0 / 1 to bf16 rounding (relative RMS <= 3e-3 at 2916 keys), not bit for bit. */
int sr_set_attention_lean(int variant);
/* Diagnostics (tools/attn_trace.py): device buffer of 4 x 16 x 16 int64 that receives clock64 stamps of the first 16 work items of
 * CTAs 0..3 of the variant-2 kernel (slots 0-7: a softmax thread, 8-11: the MMA thread); NULL (default) switches it off. */
int sr_set_attention_trace(void* buffer);

/* ---- attention (tcgen05 + TMA) --------------------------------------------------------------
 * Variable-length softmax(Q K^T * scale) V; query row r attends key rows k with row_lo[r] <= k < row_hi[r].
 * Replaces flash_attn_varlen_func / eager per-window attention at HF:244-283 (ViT windowed + full attention,
 * cu_window_seqlens / cu_seqlens) and the causal GQA prefill attention at HF:740-756.
 * Q [n_q_rows, n_q_heads, 128], K [n_k_rows, n_kv_heads, 128] (head_dim 80 zero-padded to 128), strides in elements;
 * Vt [n_kv_heads*head_dim_v, n_vt_cols] = V transposed (row stride vt_row_stride); O [n_q_rows, n_q_heads, head_dim_v].
 * tiles: int32[n_tiles][4] = {q_row0, q_nrows(<=128), kv_lo, kv_hi}; row_lo/row_hi: int32[n_q_rows].
 * head_dim_v: 80 (ViT) or 128 (decoder). */
int sr_attention_fwd(const void* Q, int64_t q_row_stride, int64_t q_head_stride, int n_q_rows, int n_q_heads,
                     const void* K, int64_t k_row_stride, int64_t k_head_stride, int n_k_rows, int n_kv_heads,
                     const void* Vt, int64_t vt_row_stride, int n_vt_cols, void* O, int64_t o_row_stride,
                     int64_t o_head_stride, const int32_t* tiles, int n_tiles, const int32_t* row_lo,
                     const int32_t* row_hi, float scale, int head_dim_v, void* stream);

/* Decode-step attention: one query token per sample against its KV cache (the reference's vLLM decode step,
 * roll/distributed/strategy/vllm_strategy.py:114-141; HF:740-756 with use_cache).  q [batch, n_q_heads*128] (ldq),
 * Kc [rows, n_kv_heads, 128], Vt [n_kv_heads*128, ldvt]; sample b attends cache rows [kv_start[b], +kv_len[b]), kv_start
 * a multiple of 8.  Split-KV: grid = (batch, kv head, ceil(max_len/128) splits); workspace = device scratch of
 * sr_decode_attention_workspace_bytes() bytes whose first batch*n_kv_heads int32 must be ZERO before the first call
 * (the kernel re-arms them). */
int64_t sr_decode_attention_workspace_bytes(int batch, int n_kv_heads, int max_len);
int sr_decode_attention(const void* q, int64_t ldq, const void* Kc, const void* Vt, int64_t ldvt, const int32_t* kv_start,
                        const int32_t* kv_len, int max_len, void* out, int64_t ldo, int batch, int n_q_heads,
                        int n_kv_heads, float scale, void* workspace, void* stream);

/* ---- HBM-bound element-wise / gather kernels ------------------------------------------------ */
/* RMSNorm, fp32 math (HF:57-71; mcore_adapter/src/mcore_adapter/models/model_utils.py:91-106). dim in {256,1280,2048}. */
int sr_rmsnorm_bf16(const void* x, int64_t ldx, const float* weight, void* y, int64_t ldy, int rows, int dim, float eps,
                    void* stream);
/* dst[i,:] = src[idx[i],:] (idx NULL = identity); src bf16 or fp32 (cast to bf16).  Window re-order and its inverse
 * (HF:478-485,511-513), last-token rows for the LM head. */
int sr_gather_rows(const void* src, int64_t lds, int src_dtype, const int32_t* idx, void* dst, int64_t ldd, int rows,
                   int cols, void* stream);
/* out[i,:] = vis_idx[i] >= 0 ? vis[vis_idx[i],:] : table[ids[i],:]  (embed_tokens + masked_scatter, HF:1299-1307;
 * mcore_adapter/.../qwen2_5_vl/modeling_qwen2_5_vl.py:294-315). */
int sr_embed_scatter(const void* table, int64_t ldt, const int64_t* ids, const void* vis, int64_t ldv,
                     const int32_t* vis_idx, void* out, int64_t ldo, int rows, int cols, void* stream);
/* ViT 2-D rotary table cos/sin [n,40] fp32 from per-patch (h,w) indices (HF:117-130,382-409). inv_freq fp32[20]. */
int sr_vit_rope_table(const int32_t* hpos, const int32_t* wpos, const float* inv_freq, float* cos_out, float* sin_out,
                      int n, void* stream);
/* Decoder M-RoPE table cos/sin [S,64] fp32 from pos3 int64[3,S] with section select (t,h,w) = (sec_t,sec_h,sec_w)
 * (mcore_adapter/.../qwen2_5_vl/modeling_qwen2_5_vl.py:21-75; HF:595-608,627-669). inv_freq fp32[64]. */
int sr_mrope_table(const int64_t* pos3, const float* inv_freq, float* cos_out, float* sin_out, int S, int sec_t,
                   int sec_h, int sec_w, void* stream);
/* ViT: split qkv [n, 3*16*80], apply 2-D RoPE (HF:156-167), write Qp/Kp [n,16,128] columns 0..79 (the caller zero-fills the
 * padding columns 80..127 once; they are never written here) and Vt [16*80, ldvt]. */
int sr_vit_rope_split(const void* qkv, int64_t ldqkv, const float* cosb, const float* sinb, void* Qp, void* Kp, void* Vt,
                      int64_t ldvt, int n, void* stream);
/* Decoder: M-RoPE on q (in place) and k; append k -> Kc[cache_row[s]], v -> Vt[:, cache_row[s]] (KV cache). */
int sr_mrope_kv(void* qkv, int64_t ldqkv, const float* cosb, const float* sinb, const int32_t* cache_row, void* Kc,
                void* Vt, int64_t ldvt, int S, int n_q_heads, int n_kv_heads, void* stream);

/* ---- LM-head tail --------------------------------------------------------------------------- */
/* Greedy token: argmax over V per row (first index on ties). logits bf16 or fp32.  Two-stage (64 slices per row);
 * workspace: rows * 512 bytes of device scratch. */
int sr_argmax_rows(const void* logits, int64_t ld, int dtype, int rows, int V, int64_t* out, void* workspace, void* stream);
/* Sampling decode with the reference's rollout parameters (temperature -> top-k -> top-p; examples/infer/rlvr_megatron.yaml:89-95,
 * roll/distributed/strategy/vllm_strategy.py:289-309): exact top-k (1 <= top_k <= 1024) by radix select, candidates ordered
 * (value desc, index asc), nucleus keeps j iff the exclusive prefix mass < top_p, inverse-CDF pick with the caller's uniform
 * uniforms[row*u_ld + *step] in [0,1) (step may be NULL = 0).  fp32 logits. */
int sr_sample_topk_topp(const float* logits, int64_t ld, int rows, int V, const float* uniforms, int u_ld, const int32_t* step,
                        float temperature, int top_k, float top_p, int64_t* out, void* stream);
/* log_probs_from_logits / entropy_from_logits (roll/utils/functionals.py:193-205). Either output may be NULL. */
int sr_logprob_entropy(const void* logits, int64_t ld, int dtype, int rows, int V, const int64_t* labels, float* logprob,
                       float* entropy, void* stream);

/* Fused LM head + log-prob / entropy for compute_log_probs (roll/pipeline/base_worker.py:206-246 -> forward_func_log_probs ->
 * strategy.py:123-138 -> roll/utils/functionals.py:193-205; lm_head HF:1385): logits = A[M,K] . W[N,K]^T are reduced to
 * per-row softmax statistics inside the tcgen05 GEMM epilogue and merged by a second kernel, so the [M, N] fp32 logits
 * (0.6 MB per position for N = 151936) never reach HBM.  logprob[r] = log_softmax(logits[r])[labels[r]], entropy[r] =
 * logsumexp(logits[r]) - sum softmax(logits[r]) * logits[r]; either output may be NULL.  labels: int64 [M] (a label outside
 * [0, N) yields logprob = -logsumexp).  workspace: sr_lmhead_logprob_workspace_bytes(M, N) bytes, 16-byte aligned. */
int64_t sr_lmhead_logprob_workspace_bytes(int M, int N);
int sr_lmhead_logprob_entropy(const void* A, int64_t lda, const void* W, int64_t ldw, const int64_t* labels, int M, int N, int K,
                              void* workspace, float* logprob, float* entropy, void* stream);

/* Decode-loop bookkeeping on the device (makes one decode step a replayable CUDA graph): tokens[:, *step] = next_tokens;
 * cache_row = row0 + *step; kv_len = len0 + *step + 1; pos3[3,batch] = pos0 + 1 + *step; (*step)++.
 * Response positions follow roll/utils/functionals.py:838-852. */
int sr_decode_advance(const int64_t* next_tokens, int64_t* tokens, int T, int32_t* step, const int32_t* row0,
                      const int32_t* len0, const int64_t* pos0, int32_t* cache_row, int32_t* kv_len, int64_t* pos3,
                      int batch, void* stream);

/* ---- persistent decode step ------------------------------------------------------------------
 * ONE launch for a whole decode step of M = batch <= 8 live sequences: all decoder layers (RMSNorm + q/k/v, M-RoPE + KV-cache
 * append + attention over the cache, o-proj + residual, RMSNorm + gate/up + SwiGLU, down + residual; HF:762-827, 672-759,
 * 611-624 with use_cache) + final norm + lm_head (HF:1385) -> fp32 logits.  Replaces the body of the reference's rollout
 * decode loop (vLLM engine step, roll/distributed/strategy/vllm_strategy.py:114-141).  Persistent kernel, one CTA per SM
 * (cooperative launch: the CTAs wait for each other's data): a producer warp streams the packed weights through a
 * shared-memory ring with cp.async.bulk without ever draining between layers, consumer warps feed mma.sync.  There is no grid
 * barrier: activations cross CTAs in flag-carrying 16-byte lines {data, gen, data, gen} that the reader polls (one L2 round
 * trip per exchange).  HBM-bound (all decoder weights + LM head per launch).
 *   layer_table: DEVICE array of n_layers sr_decode_layer (weights in the sr_gemm_skinny_packed_bf16 streaming layout).
 *   x [batch, 2048] bf16 (ldx): in = embedded tokens (read only, unless the dump flag is set).
 *   qkv/ao/act scratch: bf16 [batch, (n_q+2 n_kv)*128], [batch, n_q*128], [batch, intermediate] (dense rows); written ONLY
 *   when tuning flag bit 1 (dump) is set -- then they and x receive plain copies of the last layer's intermediates (bring-up /
 *   tests); may be NULL otherwise.
 *   cosb/sinb fp32 [batch, 64] (sr_mrope_table of this step's positions); cache_row/kv_start/kv_len as sr_decode_advance
 *   writes them (kv_len INCLUDES the new token, cache_row = kv_start + kv_len - 1).
 *   Kc [n_layers][rows, n_kv, 128], Vt [n_layers][n_kv*128, ldvt]: layer strides in ELEMENTS.
 *   workspace: sr_decode_persistent_workspace_bytes() bytes, 256-byte aligned, ZEROED once by the caller (it holds the
 *   generation counter and the flag-carrying exchange buffers; launches on one workspace must be stream-ordered).
 *   stop_after_phases: 0 = everything; k > 0 runs only the first k phases (5 per layer: qkv, attention, o-proj, gate/up,
 *   down; then the LM head) -- bring-up / test aid.
 * Hidden size is fixed at 2048 (= n_q_heads * 128); intermediate % 32 == 0; vocab % 16 == 0. */
typedef struct sr_decode_layer {
  const void* qkv_wp;   /* packed [ (n_q+2 n_kv)*128, 2048 ] */
  const float* qkv_b;   /* [(n_q+2 n_kv)*128] or NULL */
  const float* ln1;     /* input_layernorm weight [2048] */
  const void* o_wp;     /* packed [2048, 2048] */
  const float* ln2;     /* post_attention_layernorm weight [2048] */
  const void* gu_wp;    /* packed [2*intermediate, 2048], rows interleaved 16 gate | 16 up */
  const void* down_wk;  /* K-sliced [intermediate/16][4][8][4][16 rows][16 k]: socioreasoner_b200.weights._kslice_pack */
  const void* reserved; /* NULL */
} sr_decode_layer;
int64_t sr_decode_persistent_workspace_bytes(int n_q_heads, int n_kv_heads, int intermediate);
/* Tuning knobs (none changes results; negative = keep).  flags (env SR_DECODE_FLAGS, default 17): bit 0 = L2 evict_first hint on
 * the weight copies, bit 1 = dump plain intermediates into x / the scratch buffers, bit 4 = L2 prefetch of each layer's K/V slice at
 * the start of the layer; bits 2, 3, 5 are measurement aids (weight stream gated off during the exchange chain; K/V requested only
 * after q arrived; LM head on every second CTA).  ring_stages: 16 KB shared-memory ring stages in use, 2..10 (0 = default 8; env
 * SR_DECODE_RING).  off_*: unit -> CTA rotation per matrix (unit u runs on CTA (u + offset) mod #SMs; default 96, 40, 0, 20; env
 * SR_DECODE_UNIT_OFFSETS="q,o,g,d"). */
int sr_set_decode_persistent_tuning(int flags, int ring_stages, int off_qkv, int off_o, int off_gate_up, int off_down);
/* Tuning aid: when buffer != NULL every later sr_decode_step_persistent launch records (globaltimer ns, clock64) pairs at its
 * phase boundaries into buffer[n_SMs][2 roles: producer, consumer][events_per_role][2] (uint64); NULL switches it off. */
int sr_set_decode_persistent_trace(void* buffer, int events_per_role);
int sr_decode_step_persistent(const void* layer_table, int n_layers, const void* lm_head_packed, const float* final_norm,
                              int vocab, void* x, int64_t ldx, void* qkv_scratch, void* ao_scratch, void* act_scratch,
                              float* logits, int64_t ld_logits, const float* cosb, const float* sinb, const int32_t* cache_row,
                              const int32_t* kv_start, const int32_t* kv_len, void* Kc, int64_t k_layer_stride, void* Vt,
                              int64_t vt_layer_stride, int64_t ldvt, int batch, int n_q_heads, int n_kv_heads, int intermediate,
                              float norm_eps, float scale, void* workspace, int stop_after_phases, void* stream);

/* ---- image preprocessing --------------------------------------------------------------------
 * uint8 RGB tile [height, width, 3] (HWC, the PIL layout) -> pixel_values rows [N, 6 * patch^2] (N = out_height/patch * out_width/patch):
 * PIL's 8-bit bicubic resize (process_image, roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py:126-144) + the HF processor the
 * collator calls (roll/datasets/collator.py:444-565 -> Qwen2VLImageProcessor: rescale, normalise, T = 2 duplicate, merge-block
 * patchify), bit-exact with that CPU path.  h_/v_bounds int32 [out, 2] = (first input index, taps), h_/v_coeffs int32 [out, ksize]:
 * PIL's 22-bit fixed-point tables for (width -> out_width) / (height -> out_height) (socioreasoner_b200.preprocess.pil_bicubic_coeffs);
 * NULL when that axis keeps its size.  scratch: uint8 [height, out_width, 3] (only when the width changes).  mean3 / std3: HOST
 * float[3], the processor's fused constants (mean * 255, std * 255 in fp32).  out_dtype SR_DTYPE_F32 | SR_DTYPE_BF16; row_map
 * (optional, device int32 [N]): destination row of processor-order row r (lets the ViT's window permutation be applied here). */
int sr_preprocess_tile(const uint8_t* image, int height, int width, int out_height, int out_width, const int32_t* h_bounds,
                       const int32_t* h_coeffs, int h_ksize, const int32_t* v_bounds, const int32_t* v_coeffs, int v_ksize,
                       uint8_t* scratch, const float* mean3, const float* std3, void* pixel_values, int out_dtype, int64_t ld_out,
                       const int32_t* row_map, int patch_size, void* stream);

/* ---- SAM2 image path -------------------------------------------------------
 * What SAM2ImagePredictor.set_image / predict compute behind roll/distributed/strategy/seg_strategy.py:47-58 (the external ``sam2``
 * package loaded by roll/models/model_providers.py:540-545; arithmetic pinned on transformers/models/sam2/modeling_sam2.py).  The
 * linear layers and the Hiera window / global attention run on sr_gemm_bf16 / sr_attention_fwd; these are the remaining pieces. */
/* uint8 HWC image -> fp32 [3, size, size]: /255, bilinear resize (align_corners = false), (v - mean) / std; mean3 / std3 HOST float[3]. */
int sr_sam2_preprocess(const uint8_t* image, int height, int width, int size, const float* mean3, const float* std3, float* out,
                       void* stream);
/* Sam2PatchEmbeddings + windowed positional embedding: Conv2d(3 -> channels, kernel, stride, padding) of fp32 [3,size,size] + bias +
 * pos_table[row], written as bf16 rows at dst_rows[y * R + x] (window-major token order).  weight fp32 [channels, 3*kernel^2]. */
int sr_sam2_patch_embed(const float* pixels, int size, int channels, int kernel, int stride, int padding, const float* weight,
                        const float* bias, const float* pos_table, const int32_t* dst_rows, void* out, void* stream);
/* nn.LayerNorm over the last dimension (Sam2MultiScaleBlock.layer_norm1/2, the two-way transformer's norms): bf16 rows, fp32 params. */
int sr_layernorm_bf16(const void* x, int64_t ldx, int rows, int cols, const float* weight, const float* bias, float eps, void* out,
                      int64_t ldo, void* stream);
/* do_pool (2x2 max-pool of queries / residual at Hiera stage changes): out[j] = max over rows idx4[j][0..3] of x. */
int sr_pool4_rows(const void* x, int64_t ldx, const int32_t* idx4, int n_out, int cols, void* out, void* stream);
/* Sam2VisionNeck top-down path: out[(y,x)] = fine[(y,x)] + coarse[(y/2, x/2)] (nearest 2x), rows in (y, x) order. */
int sr_upsample2x_add(const void* fine, const void* coarse, int coarse_res, int cols, void* out, void* stream);
/* op 0: out = a + b (bf16); 1: relu in place on a (bf16); 2: sigmoid in place on a (fp32).  n elements. */
int sr_elementwise(int op, void* a, const void* b, void* out, int64_t n, void* stream);
/* Sam2PromptEncoder point / box-corner embedding: coords fp32 [n,2] (model-input pixels), labels int32 [n] (1/0 point, 2/3 box corner,
 * -1 "not a point"); gauss fp32 [2, hidden/2], point_embed fp32 [4, hidden], not_a_point fp32 [hidden] -> bf16 [n, hidden]. */
int sr_sam2_prompt_embed(const float* coords, const int32_t* labels, int n_points, const float* gauss, const float* point_embed,
                         const float* not_a_point, int hidden, float image_size, void* out, void* stream);
/* Sam2Attention core for the mask decoder (token <-> image, head_dim 16 or 32): softmax(q k^T / sqrt(d)) v, rows [*, heads*head_dim]. */
int sr_small_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int n_q, int n_k, int heads,
                       int head_dim, void* out, int64_t ldo, void* stream);
/* Mask-decoder upscaling step: y [in_res^2, 4*channels] = GEMM form of ConvTranspose2d(k=2,s=2) (columns dy,dx,c) -> pixel shuffle,
 * + bias + feat [(2 in_res)^2, channels], optional channel LayerNorm (ln_weight != NULL), exact GELU -> bf16 [(2 in_res)^2, channels]. */
int sr_sam2_upscale(const void* y, const void* feat, const float* bias, const float* ln_weight, const float* ln_bias, float eps, int in_res,
                    int channels, void* out, void* stream);
/* masks = hyper_in @ upscaled^T: out fp32 [count, n_pixels] from hyper bf16 [*, channels] rows first..first+count, upscaled bf16 [n_pixels, channels]. */
int sr_sam2_mask_logits(const void* hyper, const void* upscaled, int first, int count, int channels, int64_t n_pixels, float* out,
                        void* stream);

/* ---- mask epilogue + IoU -------------------------------------------------------------------- */
/* seg_strategy.py:47-65: per object pick argmax(scores) of 3 candidates, bilinear up-sample h x w -> image_size^2
 * (align_corners=False), > 0, OR over objects, cv2.INTER_NEAREST resize to out_size^2, uint8 {0,1}. */
int sr_mask_epilogue(const float* low_res_logits, const float* scores, int n_obj, int h, int w, int image_size,
                     int out_size, uint8_t* out_mask, void* stream);
/* compute_giou (rlvr_socioseg_vlm_pipeline_infer.py:45-58): iou[i] = |p&g| / |p|g| on >0 masks, empty union ->
 * empty_value (1.0 infer pipeline, 0.0 reward worker socioseg_rule_reward_worker.py:181-199). counts (optional)
 * int32[n_masks][2] = {inter, union}. */
int sr_mask_iou(const uint8_t* pred, const uint8_t* gt, int n_masks, int64_t n_pix, float empty_value, float* iou,
                int32_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SOCIOREASONER_B200_H */
