"""Qwen2.5-VL forward on the hand-written sm_100a kernels: ViT -> patch merger -> embed scatter -> decoder -> logits,
greedy KV-cache decode, mask epilogue + IoU.

Follows (arithmetic only; nothing is imported from there) HF ``transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py``
-- ``Qwen2_5_VisionTransformerPretrainedModel.forward`` (:455-518), ``Qwen2_5_VLVisionBlock`` (:290-321),
``Qwen2_5_VLPatchMerger`` (:133-146), ``Qwen2_5_VLModel.forward`` (:1270-1345), ``Qwen2_5_VLDecoderLayer`` (:762-827) --
which is what the reference's eager strategy executes (``roll/distributed/strategy/hf_strategy.py:84-90``).

Internally sequences are PACKED (padding dropped, varlen attention); the padded [B,S] API of the reference is kept at
the surface.  All device work goes through ``lib`` (C ABI); torch only allocates buffers and moves bytes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib, plan
from .config import SocioReasonerConfig
from .weights import PackedWeights


def _dev(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


@dataclass
class DevAttnPlan:
    tiles: torch.Tensor
    row_lo: torch.Tensor
    row_hi: torch.Tensor

    @staticmethod
    def make(p: plan.AttnPlan, device) -> "DevAttnPlan":
        return DevAttnPlan(_dev(p.tiles, device), _dev(p.row_lo, device), _dev(p.row_hi, device))


@dataclass
class DevVitPlan:
    host: plan.VitPlan
    row_perm: torch.Tensor
    hpos: torch.Tensor
    wpos: torch.Tensor
    window: DevAttnPlan
    full: DevAttnPlan
    cos: Optional[torch.Tensor] = None
    sin: Optional[torch.Tensor] = None


@dataclass
class SeqPlan:
    """Packed view of a padded [B,S] batch."""
    B: int
    S: int
    lengths: np.ndarray          # [B]
    q_starts: np.ndarray         # [B] packed row of each sample's first token
    flat_index: np.ndarray       # [T] index into the flattened [B*S] grid of each packed token
    total: int
    last_pos: Optional[np.ndarray] = None  # [B] max M-RoPE position over the valid tokens of each sample


class KVCache:
    def __init__(self, n_layers: int, batch: int, l_max: int, n_kv: int, head_dim: int, device):
        self.batch, self.l_max, self.n_kv, self.head_dim = batch, l_max, n_kv, head_dim
        rows = batch * l_max
        self.k = torch.zeros((n_layers, rows, n_kv, head_dim), dtype=torch.bfloat16, device=device)
        self.vt = torch.zeros((n_layers, n_kv * head_dim, rows), dtype=torch.bfloat16, device=device)
        self.lengths = np.zeros(batch, dtype=np.int64)

    @property
    def kv_starts(self) -> np.ndarray:
        return np.arange(self.batch, dtype=np.int64) * self.l_max


class PendingDecode:
    """The decode half of ``Engine.generate_greedy`` on the CUDA-graph path: ViT + prefill have been submitted, the first token's
    logits are in ``st['logits']`` and the captured decode step can be replayed ``T - 1`` times.  ``advance(n)`` submits up to n
    replays (a pipelined caller interleaves them with the next batch's prefill, which must run on ANOTHER ``Engine`` over the same
    weights: buffers, KV cache and decode state belong to the engine); ``finish()`` submits the rest and returns the tokens."""

    def __init__(self, engine, st, cache, sp, graph, T, sampling, pad_id, check_every, ev):
        self.engine, self.st, self.cache, self.sp, self.graph, self.T = engine, st, cache, sp, graph, T
        self.sampling, self.pad_id, self.check_every, self.ev = sampling, pad_id, check_every, ev
        self.done_steps, self.stopped = 0, False

    @property
    def remaining(self) -> int:
        return 0 if self.stopped else self.T - 1 - self.done_steps

    def advance(self, n: int) -> int:
        st, run = self.st, 0
        while run < n and self.remaining > 0:
            self.graph.replay()
            run += 1
            self.done_steps += 1
            if st["track_eos"] and self.done_steps % self.check_every == 0 and bool(st["done"].all().item()):
                self.stopped = True
        lib.launch_count += run * self.engine._decode_launches
        return run

    def ticker(self, n_ticks: int, first: int = 0, last: Optional[int] = None):
        """A ``tick`` callback that spreads the remaining decode steps evenly over calls ``first`` .. ``last`` - 1 of ``n_ticks``
        (default: all of them); calls outside that window submit nothing, the last call of the window completes the count."""
        last = n_ticks if last is None else last
        span = max(last - first, 1)
        total, state = self.remaining, {"i": 0, "done": 0}

        def tick():
            state["i"] += 1
            k = min(max(state["i"] - first, 0), span)
            want = (total * k + span - 1) // span
            if want > state["done"]:
                state["done"] += self.advance(want - state["done"])
        return tick

    def finish(self) -> torch.Tensor:
        eng, st = self.engine, self.st
        self.advance(self.remaining)
        if not self.stopped:  # the last token: pick + append (no further forward needed)
            eng._pick_token(st, self.sampling)
            lib.decode_advance(st["nxt"], st["tokens"], st["step"], st["row0"], st["len0"], st["pos0"], st["cache_row"],
                               st["kv_len"], st["pos3"])
        self.cache.lengths = self.sp.lengths + self.T - 1
        if self.ev is not None:
            self.ev.append(("decode", eng._event()))
        tokens = st["tokens"].clone()
        if st["track_eos"]:  # everything after a sequence's first stop id is padding
            is_eos = (tokens.unsqueeze(-1) == st["eos"].view(1, 1, -1)).any(-1)
            after = (is_eos.cumsum(1) - is_eos.long()) > 0
            tokens = torch.where(after, torch.full_like(tokens, self.pad_id), tokens)
        return tokens


class Engine:
    def __init__(self, cfg: SocioReasonerConfig, weights: PackedWeights, device="cuda"):
        cfg.validate()
        lib.load()
        self.cfg, self.w, self.device = cfg, weights, torch.device(device)
        v, t = cfg.vision, cfg.text
        # inv_freq with the reference's exact fp32 expressions (HF:122, mcore_adapter .../modeling_qwen2_5_vl.py:62)
        d = v.head_dim // 2
        self.vit_inv_freq = (1.0 / (v.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float) / d))).to(self.device)
        self.txt_inv_freq = (1.0 / (t.rope_theta ** (torch.arange(0, t.head_dim, 2, dtype=torch.float32) / t.head_dim))).to(self.device)
        self._bufs: Dict[str, torch.Tensor] = {}
        self._vit_plans: Dict[bytes, DevVitPlan] = {}
        self._caches: Dict[Tuple[int, int], KVCache] = {}
        self._graphs: Dict[Tuple[int, int, int], "torch.cuda.CUDAGraph"] = {}
        self._dstate: Dict[Tuple[int, int], dict] = {}
        self._decode_launches = 0
        self.use_cuda_graph = True
        self.vit_narrow_block_n = int(os.environ.get("SR_VIT_N1280_BLOCK_N", "0"))  # 0 = auto (256)
        # persistent decode step (sr_decode_persistent.cu): one launch per token for <= 8 live sequences; SR_DECODE_PERSISTENT=0
        # selects the per-layer kernels (the only path for larger batches)
        self.use_persistent_decode = os.environ.get("SR_DECODE_PERSISTENT", "1") != "0"
        self._layer_table: Optional[torch.Tensor] = None
        self._sm_count = None
        self.stage_events = None  # optional list: generate_greedy appends (stage, cuda event) pairs (bench.py stage split)

    @staticmethod
    def _event():
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        return e

    # ------------------------------------------------------------------ buffers
    def _buf(self, name: str, shape, dtype=torch.bfloat16, zero: bool = False) -> torch.Tensor:
        n = int(np.prod(shape))
        cur = self._bufs.get(name)
        if cur is None or cur.dtype != dtype or cur.numel() < n:
            cur = (torch.zeros if zero else torch.empty)(max(n, 1), dtype=dtype, device=self.device)
            self._bufs[name] = cur
        return cur[:n].view(*shape)

    def free_buffers(self) -> None:
        self._graphs.clear()
        self._bufs.clear()
        self._vit_plans.clear()
        self._caches.clear()
        self._dstate.clear()

    # ------------------------------------------------------------------ ViT
    def vit_plan(self, grid_thw) -> DevVitPlan:
        g = np.ascontiguousarray(np.asarray(grid_thw.cpu() if torch.is_tensor(grid_thw) else grid_thw, dtype=np.int64))
        key = g.tobytes()
        p = self._vit_plans.get(key)
        if p is None:
            v = self.cfg.vision
            hp = plan.vit_plan(g, v.spatial_merge_size, v.window_size, v.patch_size)
            p = DevVitPlan(hp, _dev(hp.row_perm, self.device), _dev(hp.hpos, self.device), _dev(hp.wpos, self.device),
                           DevAttnPlan.make(hp.window_plan, self.device), DevAttnPlan.make(hp.full_plan, self.device))
            p.cos, p.sin = lib.vit_rope_table(p.hpos, p.wpos, self.vit_inv_freq)
            if len(self._vit_plans) > 64:
                self._vit_plans.clear()
            self._vit_plans[key] = p
        return p

    def vit(self, pixel_values: torch.Tensor, grid_thw, capture: Optional[List[torch.Tensor]] = None, tick=None
            ) -> Tuple[torch.Tensor, DevVitPlan]:
        """pixel_values [N,1176] (fp32 or bf16, processor row order) -> merged embeddings [M, 2048] in WINDOW order
        (``vp.host.merged_reverse`` maps original order -> window order) and the plan."""
        v, W = self.cfg.vision, self.w
        vp = self.vit_plan(grid_thw)
        n = vp.host.n
        assert pixel_values.shape == (n, v.patch_dim), (pixel_values.shape, n)
        D, H, hd = v.hidden_size, v.num_heads, v.head_dim
        x0 = lib.gather_rows(pixel_values, vp.row_perm, out=self._buf("vit_x0", (n, v.patch_dim)))
        h = lib.gemm(x0, W.patch_w, out=self._buf("vit_h", (n, D)))
        y = self._buf("vit_y", (n, D))
        qkv = self._buf("vit_qkv", (n, 3 * D))
        qp = self._buf("vit_qp", (n, H, 128), zero=True)  # head_dim 80 -> 128: padding columns stay zero (never written)
        kp = self._buf("vit_kp", (n, H, 128), zero=True)
        npad = (n + 7) // 8 * 8
        vt = self._buf("vit_vt", (H * hd, npad), zero=True)
        ao = self._buf("vit_ao", (n, H, hd))
        act = self._buf("vit_act", (n, v.intermediate_padded))
        scale = hd ** -0.5
        bn_narrow = self.vit_narrow_block_n  # tile width for the N = 1280 GEMMs (915 tiles of 128x256 = 6.2 waves on 148 SMs)
        for i, b in enumerate(W.vit_blocks):
            lib.rmsnorm(h, b.norm1, v.rms_norm_eps, out=y)
            lib.gemm(y, b.qkv_w, out=qkv, bias=b.qkv_b)
            lib.vit_rope_split(qkv, vp.cos, vp.sin, qp, kp, vt)
            ap = vp.full if i in v.fullatt_block_indexes else vp.window
            lib.attention(qp, kp, vt, ao, ap.tiles, ap.row_lo, ap.row_hi, scale, hd)
            lib.gemm(ao.view(n, D), b.proj_w, out=h, bias=b.proj_b, residual=h, block_n=bn_narrow)
            lib.rmsnorm(h, b.norm2, v.rms_norm_eps, out=y)
            lib.gemm(y, b.gu_w, out=act, bias=b.gu_b, epilogue=lib.EPI_SWIGLU)
            lib.gemm(act, b.down_w, out=h, bias=b.down_b, residual=h, block_n=bn_narrow)
            if capture is not None:
                capture.append(h.clone())
            if tick is not None:
                tick()
        lib.rmsnorm(h, W.merger_ln, v.rms_norm_eps, out=y)
        m = vp.host.m
        hm = D * v.spatial_merge_size ** 2
        z = lib.gemm(y.view(m, hm), W.merger_w0, out=self._buf("vit_z", (m, hm)), bias=W.merger_b0, epilogue=lib.EPI_GELU)
        merged = lib.gemm(z, W.merger_w2, out=self._buf("vit_merged", (m, v.out_hidden_size)), bias=W.merger_b2)
        return merged, vp

    def image_features(self, pixel_values, grid_thw) -> torch.Tensor:
        """Merged embeddings in ORIGINAL order (what HF's visual(...).pooler_output returns, HF:511-518)."""
        merged, vp = self.vit(pixel_values, grid_thw)
        rev = _dev(vp.host.merged_reverse, self.device)
        return lib.gather_rows(merged, rev)

    # ------------------------------------------------------------------ sequence packing
    @staticmethod
    def seq_plan(attention_mask: np.ndarray) -> SeqPlan:
        B, S = attention_mask.shape
        keep = attention_mask.reshape(-1) != 0
        flat = np.nonzero(keep)[0]
        lengths = (attention_mask != 0).sum(axis=1).astype(np.int64)
        q_starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
        return SeqPlan(B, S, lengths, q_starts, flat, int(lengths.sum()))

    # ------------------------------------------------------------------ decoder
    def _decoder_layers(self, x: torch.Tensor, cos, sin, cache: KVCache, cache_row: torch.Tensor,
                        attn_plan: Optional[DevAttnPlan], decode: Optional[Tuple[torch.Tensor, torch.Tensor, int]],
                        capture: Optional[List[torch.Tensor]] = None, tick=None) -> torch.Tensor:
        t, W = self.cfg.text, self.w
        S = x.shape[0]
        Hq, Hkv, hd = t.num_attention_heads, t.num_key_value_heads, t.head_dim
        qkv_dim = (Hq + 2 * Hkv) * hd
        tag = f"d{S}" if decode is not None else "p"  # decode buffers are baked into CUDA graphs: never resized
        y = self._buf(tag + "_y", (S, t.hidden_size))
        qkv = self._buf(tag + "_qkv", (S, qkv_dim))
        ao = self._buf(tag + "_ao", (S, Hq * hd))
        act = self._buf(tag + "_act", (S, t.intermediate_size))
        scale = hd ** -0.5
        fuse_norm = decode is not None and S <= 32  # decode: RMSNorm folded into the weight-streaming GEMM (2 launches/layer less)
        for li, L in enumerate(W.layers):
            if fuse_norm and L.qkv_wp is not None:
                lib.gemm_skinny_packed(x, L.qkv_wp, qkv_dim, t.hidden_size, out=qkv, bias=L.qkv_b, norm_weight=L.ln1,
                                       norm_eps=t.rms_norm_eps)
            elif fuse_norm:
                lib.gemm(x, L.qkv_w, out=qkv, bias=L.qkv_b, norm_weight=L.ln1, norm_eps=t.rms_norm_eps)
            else:
                lib.rmsnorm(x, L.ln1, t.rms_norm_eps, out=y)
                lib.gemm(y, L.qkv_w, out=qkv, bias=L.qkv_b)
            lib.mrope_kv(qkv, cos, sin, cache_row, cache.k[li], cache.vt[li], Hq, Hkv)
            if decode is None:
                q = qkv[:, : Hq * hd].view(S, Hq, hd)
                lib.attention(q, cache.k[li], cache.vt[li], ao.view(S, Hq, hd), attn_plan.tiles, attn_plan.row_lo,
                              attn_plan.row_hi, scale, hd)
            else:
                kv_start, kv_len, max_len = decode
                lib.decode_attention(qkv, cache.k[li], cache.vt[li], kv_start, kv_len, max_len, ao, Hq, Hkv, scale)
            if fuse_norm and L.o_wp is not None:
                lib.gemm_skinny_packed(ao, L.o_wp, t.hidden_size, Hq * hd, out=x, residual=x)
                lib.gemm_skinny_packed(x, L.gu_wp, 2 * t.intermediate_size, t.hidden_size, out=act, epilogue=lib.EPI_SWIGLU,
                                       norm_weight=L.ln2, norm_eps=t.rms_norm_eps)
                lib.gemm_skinny_packed(act, L.down_wp, t.hidden_size, t.intermediate_size, out=x, residual=x)
                if capture is not None:
                    capture.append(x.clone())
                continue
            lib.gemm(ao, L.o_w, out=x, residual=x)
            if fuse_norm:
                lib.gemm(x, L.gu_w, out=act, epilogue=lib.EPI_SWIGLU, norm_weight=L.ln2, norm_eps=t.rms_norm_eps)
            else:
                lib.rmsnorm(x, L.ln2, t.rms_norm_eps, out=y)
                lib.gemm(y, L.gu_w, out=act, epilogue=lib.EPI_SWIGLU)
            lib.gemm(act, L.down_w, out=x, residual=x)
            if capture is not None:
                capture.append(x.clone())
            if tick is not None:
                tick()
        return x

    def prefill_ticks(self, with_images: bool = True) -> int:
        """How often ``prefill(..., tick=f)`` calls ``f``: once per ViT block and once per decoder layer."""
        return (len(self.w.vit_blocks) if with_images else 0) + len(self.w.layers)

    def prefill(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, position_ids: Optional[torch.Tensor],
                pixel_values: Optional[torch.Tensor], image_grid_thw, cache: Optional[KVCache] = None,
                reserve_new_tokens: int = 0, capture: Optional[List[torch.Tensor]] = None, tick=None):
        """Runs ViT + embed/scatter + all decoder layers over the packed tokens.
        Returns (hidden [T, 2048] packed (pre final-norm), SeqPlan, KVCache, position_ids [3,B,S] numpy)."""
        cfg, t = self.cfg, self.cfg.text
        ids_cpu = input_ids.detach().cpu().numpy() if torch.is_tensor(input_ids) else np.asarray(input_ids)
        mask_cpu = (attention_mask.detach().cpu().numpy() if torch.is_tensor(attention_mask) else np.asarray(attention_mask)) \
            if attention_mask is not None else np.ones_like(ids_cpu)
        grid_cpu = None
        if image_grid_thw is not None:
            grid_cpu = image_grid_thw.detach().cpu().numpy() if torch.is_tensor(image_grid_thw) else np.asarray(image_grid_thw)
        sp = self.seq_plan(mask_cpu)
        if position_ids is None:
            pos3, _ = plan.rope_index(ids_cpu, grid_cpu, mask_cpu, cfg.image_token_id, cfg.vision_start_token_id,
                                      cfg.vision.spatial_merge_size)
        else:
            pos3 = position_ids.detach().cpu().numpy() if torch.is_tensor(position_ids) else np.asarray(position_ids)
            if pos3.shape[0] != 3:  # [B,3,S] as stored in DataProto (hf_strategy.py:65-67 transposes it)
                pos3 = pos3.transpose(1, 0, 2)
        sp.last_pos = np.where(mask_cpu[None] != 0, pos3, np.iinfo(np.int64).min).max(axis=(0, 2)).astype(np.int64)
        ids_packed = ids_cpu.reshape(-1)[sp.flat_index].astype(np.int64)
        pos_packed = np.ascontiguousarray(pos3.reshape(3, -1)[:, sp.flat_index]).astype(np.int64)
        # vision
        merged, vis_idx = None, None
        if pixel_values is not None:
            pv = pixel_values if pixel_values.is_cuda else pixel_values.to(self.device, non_blocking=True)
            merged, vp = self.vit(pv, grid_cpu, tick=tick)
            is_img = ids_packed == cfg.image_token_id
            n_img_tok = int(is_img.sum())
            if n_img_tok != vp.host.m:  # same check as HF get_placeholder_mask (HF:1179-1218)
                raise ValueError(f"Image features and image tokens do not match: tokens: {n_img_tok}, features {vp.host.m}")
            vi = np.full(sp.total, -1, dtype=np.int32)
            vi[is_img] = vp.host.merged_reverse
            vis_idx = _dev(vi, self.device)
        # KV cache
        if cache is None:
            l_max = int((sp.lengths.max() + reserve_new_tokens + 127) // 128 * 128)
            cache = KVCache(t.num_hidden_layers, sp.B, l_max, t.num_key_value_heads, t.head_dim, self.device)
        assert cache.batch == sp.B and sp.lengths.max() + reserve_new_tokens <= cache.l_max
        cache.lengths = sp.lengths.copy()
        kv_starts = cache.kv_starts
        cache_row_np = np.concatenate([kv_starts[b] + np.arange(sp.lengths[b]) for b in range(sp.B)]).astype(np.int32)
        ap = DevAttnPlan.make(plan.plan_causal(sp.lengths, sp.q_starts, kv_starts), self.device)
        ids_d = _dev(ids_packed, self.device)
        pos_d = _dev(pos_packed, self.device)
        cache_row = _dev(cache_row_np, self.device)
        x = lib.embed_scatter(self.w.embed, ids_d, merged, vis_idx, out=self._buf("p_x", (sp.total, t.hidden_size)))
        cos, sin = lib.mrope_table(pos_d, self.txt_inv_freq, t.mrope_section)
        x = self._decoder_layers(x, cos, sin, cache, cache_row, ap, None, capture, tick=tick)
        return x, sp, cache, pos3

    def final_logits(self, hidden_rows: torch.Tensor, out_dtype=torch.float32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """final RMSNorm + LM head on the given rows (HF:1385)."""
        t = self.cfg.text
        if hidden_rows.shape[0] <= 32:
            if self.w.lm_head_p is not None:  # separate norm: the fused-norm variant is slower on the 622 MB LM head
                y = lib.rmsnorm(hidden_rows, self.w.final_norm, t.rms_norm_eps, out=self._buf(f"ln_y{hidden_rows.shape[0]}", tuple(hidden_rows.shape)))
                return lib.gemm_skinny_packed(y, self.w.lm_head_p, t.vocab_size, t.hidden_size, out=out, out_dtype=out_dtype)
            return lib.gemm(hidden_rows, self.w.lm_head, out=out, out_dtype=out_dtype, norm_weight=self.w.final_norm,
                            norm_eps=t.rms_norm_eps)
        y = lib.rmsnorm(hidden_rows, self.w.final_norm, t.rms_norm_eps)
        return lib.gemm(y, self.w.lm_head, out=out, out_dtype=out_dtype)

    # ------------------------------------------------------------------ public forwards
    def forward_logits(self, input_ids, attention_mask, position_ids, pixel_values, image_grid_thw,
                       out_dtype=torch.float32) -> torch.Tensor:
        """Padded-API forward: logits [B,S,V]; rows at padded positions are zero."""
        x, sp, _, _ = self.prefill(input_ids, attention_mask, position_ids, pixel_values, image_grid_thw)
        V = self.cfg.text.vocab_size
        packed = self.final_logits(x, out_dtype)
        if sp.total == sp.B * sp.S:
            return packed.view(sp.B, sp.S, V)
        full = torch.zeros((sp.B * sp.S, V), dtype=out_dtype, device=self.device)
        full[_dev(sp.flat_index, self.device)] = packed
        return full.view(sp.B, sp.S, V)

    def forward_log_probs(self, input_ids, attention_mask, position_ids, pixel_values, image_grid_thw, response_mask,
                          want_entropy: bool = True):
        """``compute_log_probs`` without the logits (roll/pipeline/base_worker.py:206-246): returns (log_probs, entropy), both
        fp32 [B, S-1] with the semantics of ``op_compute_log_probs`` / ``op_compute_entropy`` (strategy.py:123-138):
        ``log_probs[b, s] = log_softmax(logits[b, s])[input_ids[b, s+1]] * response_mask[b, s+1]`` and the same mask on the
        entropy.  Positions whose mask is 0 are multiplied by 0 in the reference, so only the rows with
        ``response_mask[b, s+1] == 1`` go through the LM head -- fused with the softmax statistics in the GEMM epilogue
        (``sr_lmhead_logprob_entropy``): neither the [B, S, V] logits nor the prompt rows' logits are ever computed."""
        t = self.cfg.text
        x, sp, _, _ = self.prefill(input_ids, attention_mask, position_ids, pixel_values, image_grid_thw)
        ids = input_ids.detach().cpu().numpy() if torch.is_tensor(input_ids) else np.asarray(input_ids)
        rm = response_mask.detach().cpu().numpy() if torch.is_tensor(response_mask) else np.asarray(response_mask)
        B, S = sp.B, sp.S
        rows_np, labels_np, out_np = plan.logprob_rows(ids, rm, sp.flat_index)
        lp = torch.zeros((B, S - 1), dtype=torch.float32, device=self.device)
        ent = torch.zeros((B, S - 1), dtype=torch.float32, device=self.device) if want_entropy else None
        if len(rows_np) == 0:
            return lp, ent
        rows, labels, flat = _dev(rows_np, self.device), _dev(labels_np, self.device), _dev(out_np, self.device)
        y = lib.rmsnorm(x.index_select(0, rows), self.w.final_norm, t.rms_norm_eps)
        l_rows, e_rows = lib.lm_head_logprob_entropy(y, self.w.lm_head, labels, want_entropy=want_entropy)
        lp.view(-1).index_copy_(0, flat, l_rows)
        if want_entropy:
            ent.view(-1).index_copy_(0, flat, e_rows)
        return lp, ent

    def last_token_rows(self, x: torch.Tensor, sp: SeqPlan) -> torch.Tensor:
        idx = _dev((sp.q_starts + sp.lengths - 1).astype(np.int32), self.device)
        return lib.gather_rows(x, idx)

    def _get_cache(self, batch: int, l_max: int) -> KVCache:
        """KV caches are kept per (batch, l_max) so their addresses stay valid for captured decode graphs."""
        t = self.cfg.text
        key = (batch, l_max)
        c = self._caches.get(key)
        if c is None:
            if len(self._caches) >= 2:  # bounded: two shapes alive at most (180 GB HBM, ~40 MB / sequence-k-token)
                self._caches.clear()
                self._graphs.clear()
            c = KVCache(t.num_hidden_layers, batch, l_max, t.num_key_value_heads, t.head_dim, self.device)
            self._caches[key] = c
        return c

    def _decode_state(self, B: int, T: int):
        key = (B, T)
        st = self._dstate.get(key)
        if st is None:
            dev = self.device
            st = dict(
                tokens=torch.zeros((B, T), dtype=torch.int64, device=dev), step=torch.zeros((1,), dtype=torch.int32, device=dev),
                nxt=torch.zeros((B,), dtype=torch.int64, device=dev), row0=torch.zeros((B,), dtype=torch.int32, device=dev),
                len0=torch.zeros((B,), dtype=torch.int32, device=dev), pos0=torch.zeros((B,), dtype=torch.int64, device=dev),
                cache_row=torch.zeros((B,), dtype=torch.int32, device=dev), kv_len=torch.zeros((B,), dtype=torch.int32, device=dev),
                pos3=torch.zeros((3, B), dtype=torch.int64, device=dev), kv_start=torch.zeros((B,), dtype=torch.int32, device=dev),
                logits=torch.zeros((B, self.cfg.text.vocab_size), dtype=torch.float32, device=dev),
                u=torch.zeros((B, T), dtype=torch.float32, device=dev), forced=torch.zeros((B, T), dtype=torch.int64, device=dev),
                done=torch.zeros((B,), dtype=torch.bool, device=dev), eos=torch.full((16,), -1, dtype=torch.int64, device=dev),
                x=torch.zeros((B, self.cfg.text.hidden_size), dtype=torch.bfloat16, device=dev),
            )
            if self._persistent_ok(B):
                t = self.cfg.text
                st.update(
                    qkv=torch.zeros((B, (t.num_attention_heads + 2 * t.num_key_value_heads) * t.head_dim), dtype=torch.bfloat16, device=dev),
                    ao=torch.zeros((B, t.num_attention_heads * t.head_dim), dtype=torch.bfloat16, device=dev),
                    act=torch.zeros((B, t.intermediate_size), dtype=torch.bfloat16, device=dev),
                    ws=lib.decode_persistent_workspace(t.num_attention_heads, t.num_key_value_heads, t.intermediate_size, dev),
                )
            self._dstate[key] = st
        return st

    def set_persistent_decode(self, on: bool) -> None:
        """Switch between the persistent decode-step kernel and the per-layer decode kernels (drops captured graphs)."""
        self.use_persistent_decode = bool(on)
        self._graphs.clear()
        self._dstate.clear()

    def _persistent_ok(self, B: int) -> bool:
        t = self.cfg.text
        return (self.use_persistent_decode and B <= 8 and self.w.lm_head_p is not None and len(self.w.layers) > 0
                and self.w.layers[0].qkv_wp is not None and self.w.layers[0].down_wk is not None and t.intermediate_size % 32 == 0
                and t.vocab_size % 16 == 0)

    def _pick_token(self, st, sampling) -> None:
        if st.get("use_forced"):  # teacher forcing (parity tests): token of this step from a given [B, T] table
            B = st["nxt"].shape[0]
            st["nxt"].copy_(st["forced"].gather(1, st["step"].long().view(1, 1).expand(B, 1)).view(B))
        elif sampling is None:
            lib.argmax_rows(st["logits"], out=st["nxt"])
        else:  # (temperature, top_k, top_p): uniforms[b, step] drive the inverse-CDF pick
            lib.sample_topk_topp(st["logits"], st["u"], sampling[0], sampling[1], sampling[2], step=st["step"], out=st["nxt"])

    def _decode_step(self, st, cache: KVCache, sampling=None, kv_tokens: Optional[int] = None) -> None:
        """One decode step, entirely on the device (graph-capturable): pick token -> append/advance -> embed -> 36 layers ->
        final norm + LM head -> st['logits']."""
        t = self.cfg.text
        self._pick_token(st, sampling)
        if st.get("track_eos"):  # a sequence is finished once it has emitted one of the stop ids (that token is still appended)
            st["done"].logical_or_((st["nxt"].view(-1, 1) == st["eos"].view(1, -1)).any(1))
        lib.decode_advance(st["nxt"], st["tokens"], st["step"], st["row0"], st["len0"], st["pos0"], st["cache_row"],
                           st["kv_len"], st["pos3"])
        cos, sin = lib.mrope_table(st["pos3"], self.txt_inv_freq, t.mrope_section)
        lib.embed_scatter(self.w.embed, st["nxt"], None, None, out=st["x"])
        if "ws" in st and self.use_persistent_decode:
            if self._layer_table is None:
                self._layer_table = lib.decode_layer_table(self.w.layers, self.device)
            lib.decode_step_persistent(self._layer_table, self.w.lm_head_p, self.w.final_norm, t.vocab_size, st["x"], st["qkv"],
                                       st["ao"], st["act"], st["logits"], cos, sin, st["cache_row"], st["kv_start"], st["kv_len"],
                                       cache.k, cache.vt, t.num_attention_heads, t.num_key_value_heads, t.intermediate_size,
                                       t.rms_norm_eps, t.head_dim ** -0.5, st["ws"], kv_tokens=kv_tokens)
            return
        xo = self._decoder_layers(st["x"], cos, sin, cache, st["cache_row"], None, (st["kv_start"], st["kv_len"], cache.l_max))
        self.final_logits(xo, torch.float32, out=st["logits"])

    def generate_greedy(self, input_ids, attention_mask, position_ids, pixel_values, image_grid_thw, max_new_tokens: int,
                        return_logits: bool = False, use_graph: Optional[bool] = None, sampling=None, uniforms=None,
                        forced_tokens: Optional[torch.Tensor] = None, eos_ids: Optional[Sequence[int]] = None, pad_id: int = 0,
                        check_every: int = 16, tick=None, defer_decode: bool = False):
        """Greedy decode, EOS ignored, fixed ``max_new_tokens``.  Response positions continue from the
        last prompt position + 1 on all three axes (roll/utils/functionals.py:838-852).  The decode step is captured
        once per (batch, cache) into a CUDA graph and replayed (launch-bound loop, ~300 kernels per token).
        ``sampling`` = (temperature, top_k, top_p) switches the per-step pick from argmax to the reference's sampling
        config; ``uniforms`` [B, T] in [0,1) are the random numbers consumed (one per sample and step).
        ``forced_tokens`` [B, T] (teacher forcing, parity tests): step t appends forced_tokens[:, t] instead of its own pick, so
        the logits of every step are comparable with an oracle that was fed the same tokens.
        ``eos_ids`` (the reference's stop ids, vllm_strategy.py:289-309): a sequence is finished after emitting one of them; a
        device-side flag is read back every ``check_every`` steps and the loop ends as soon as every sequence has finished (at the
        reference's 2048-token cap a fixed-length loop would decode for seconds after the last stop).  Positions after a sequence's
        stop token, and steps that never ran, hold ``pad_id``.
        ``tick``: called ``prefill_ticks()`` times during ViT + prefill (after every ViT block / decoder layer) -- the hook a pipelined
        caller uses to submit another batch's decode steps in between.  ``defer_decode``: run ViT + prefill, set the decode up and
        return a ``PendingDecode`` instead of tokens; its ``advance(n)`` / ``finish()`` replay the captured decode step (graph path
        only).  ``generate_greedy(...)`` == ``generate_greedy(..., defer_decode=True).finish()``.
        Returns new tokens [B, T] (int64, device) and optionally fp32 logits [B, T, V]."""
        t = self.cfg.text
        T = max_new_tokens
        mask_np = attention_mask.detach().cpu().numpy() if torch.is_tensor(attention_mask) else np.asarray(attention_mask)
        lens = (mask_np != 0).sum(axis=1)
        B = mask_np.shape[0]
        l_max = int((lens.max() + T + 127) // 128 * 128)
        cache = self._get_cache(B, l_max)
        ev = self.stage_events
        if ev is not None:
            ev.append(("start", self._event()))
        x, sp, cache, pos3 = self.prefill(input_ids, attention_mask, position_ids, pixel_values, image_grid_thw, cache=cache,
                                          reserve_new_tokens=T, tick=tick)
        if ev is not None:
            ev.append(("vit+prefill", self._event()))
        V = t.vocab_size
        st = self._decode_state(B, T)
        st["step"].zero_()
        st["use_forced"] = forced_tokens is not None
        st["track_eos"] = eos_ids is not None and len(eos_ids) > 0
        if st["track_eos"]:
            assert len(eos_ids) <= 16, "at most 16 stop ids"
            st["eos"].fill_(-1)
            st["eos"][:len(eos_ids)].copy_(torch.tensor(list(eos_ids), dtype=torch.int64), non_blocking=True)
            st["done"].zero_()
            st["tokens"].fill_(int(pad_id))
        if forced_tokens is not None:
            assert tuple(forced_tokens.shape) == (B, T)
            st["forced"].copy_(forced_tokens.to(self.device, torch.int64), non_blocking=True)
        if sampling is not None:
            sampling = (float(sampling[0]), int(sampling[1]), float(sampling[2]))
            assert uniforms is not None and tuple(uniforms.shape) == (B, T)
            st["u"].copy_(uniforms.to(self.device, torch.float32), non_blocking=True)
        st["kv_start"].copy_(torch.from_numpy(cache.kv_starts.astype(np.int32)), non_blocking=True)
        st["row0"].copy_(torch.from_numpy((cache.kv_starts + sp.lengths).astype(np.int32)), non_blocking=True)
        st["len0"].copy_(torch.from_numpy(sp.lengths.astype(np.int32)), non_blocking=True)
        st["pos0"].copy_(torch.from_numpy(sp.last_pos.astype(np.int64)), non_blocking=True)
        self.final_logits(self.last_token_rows(x, sp), torch.float32, out=st["logits"])
        all_logits = torch.empty((B, T, V), dtype=torch.float32, device=self.device) if return_logits else None
        stopped = False
        if use_graph is None:
            use_graph = self.use_cuda_graph and not return_logits and lib.timeline is None
        if use_graph and T > 1:
            gkey = (B, T, l_max, sampling, self.use_persistent_decode, forced_tokens is not None, st["track_eos"])
            g = self._graphs.get(gkey)
            if g is None:
                # warm-up outside capture (buffer allocation, function attributes), then restore the state it advanced
                keep = {k: st[k].clone() for k in ("logits", "step")}
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                l0 = lib.launch_count
                with torch.cuda.stream(side):
                    self._decode_step(st, cache, sampling)
                self._decode_launches = lib.launch_count - l0
                torch.cuda.current_stream().wait_stream(side)
                st["logits"].copy_(keep["logits"]); st["step"].copy_(keep["step"])
                g = torch.cuda.CUDAGraph()
                l0 = lib.launch_count
                with torch.cuda.graph(g):
                    self._decode_step(st, cache, sampling)
                lib.launch_count = l0  # capture launches nothing
                st["logits"].copy_(keep["logits"]); st["step"].copy_(keep["step"])
                self._graphs[gkey] = g
            pend = PendingDecode(self, st, cache, sp, g, T, sampling, int(pad_id), check_every, ev)
            if defer_decode:
                return pend
            return pend.finish()
        else:
            assert not defer_decode, "defer_decode needs the CUDA-graph decode path"
            for step in range(T - 1):
                if return_logits:
                    all_logits[:, step] = st["logits"]
                l0 = lib.launch_count
                self._decode_step(st, cache, sampling, kv_tokens=int(sp.lengths.sum()) + B * (step + 1))
                self._decode_launches = lib.launch_count - l0
                if st["track_eos"] and (step + 1) % check_every == 0 and bool(st["done"].all().item()):
                    stopped = True
                    break
        if return_logits and not stopped:
            all_logits[:, T - 1] = st["logits"]
        if not stopped:  # the last token: pick + append (no further forward needed)
            self._pick_token(st, sampling)
            lib.decode_advance(st["nxt"], st["tokens"], st["step"], st["row0"], st["len0"], st["pos0"], st["cache_row"],
                               st["kv_len"], st["pos3"])
        cache.lengths = sp.lengths + T - 1
        if ev is not None:
            ev.append(("decode", self._event()))
        tokens = st["tokens"].clone()
        if st["track_eos"]:  # everything after a sequence's first stop id is padding
            is_eos = (tokens.unsqueeze(-1) == st["eos"].view(1, 1, -1)).any(-1)
            after = (is_eos.cumsum(1) - is_eos.long()) > 0
            tokens = torch.where(after, torch.full_like(tokens, int(pad_id)), tokens)
        return (tokens, all_logits) if return_logits else tokens

    # ------------------------------------------------------------------ mask epilogue
    def mask_epilogue(self, low_res_logits: torch.Tensor, scores: torch.Tensor, image_size: int = 756, out_size: int = 768,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return lib.mask_epilogue(low_res_logits, scores, image_size, out_size, out=out)

    def mask_iou(self, pred: torch.Tensor, gt: torch.Tensor, empty_value: float = 1.0) -> torch.Tensor:
        return lib.mask_iou(pred, gt, empty_value)
