"""SAM2 image forward on the hand-written kernels: Hiera trunk -> FPN neck (``set_image``), prompt encoder -> two-way mask decoder
(``predict_logits``).  Follows (arithmetic only) ``transformers/models/sam2/modeling_sam2.py`` -- ``Sam2HieraDetModel.forward``,
``Sam2MultiScaleBlock``, ``Sam2MultiScaleAttention``, ``Sam2VisionNeck``, ``Sam2PromptEncoder``, ``Sam2TwoWayTransformer``,
``Sam2MaskDecoder.forward`` with ``multimask_output=True`` -- the HF port of the ``sam2`` package the reference's
``SAM2ImagePredictor`` runs (roll/models/model_providers.py:540-545; roll/distributed/strategy/seg_strategy.py:47,58).

All device work goes through an ``ops`` object (``ops_cuda.CudaOps``: the C ABI kernels; there is no other product backend).

Both halves are launch-bound when issued from Python (hiera-large: ~430 launches per image, ~100 per object prompt, 15-20 us of host
time each), so with a backend that says ``supports_graph`` they are captured ONCE as CUDA graphs -- one per image geometry, one
per prompt length -- and replayed: the image lands in a static input buffer, the prompt in static device buffers fed from a
pinned host pair, the outputs of a prompt replay are cloned (the next replay overwrites them)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import layout
from .config import Sam2Cfg
from .weights import AttnW, Sam2Weights


class Sam2Engine:
    def __init__(self, cfg: Sam2Cfg, weights: Sam2Weights, ops):
        cfg.validate()
        self.cfg, self.w, self.ops = cfg, weights, ops
        self._plans: Dict[Tuple, object] = {}
        self.feat_s0 = self.feat_s1 = self.image_embed = None
        self.use_graph = bool(getattr(ops, "supports_graph", False))
        self._img_graph = None       # (graph, static input, (feat_s0, feat_s1, image_embed))
        self._pred_graphs: Dict[int, tuple] = {}

    # ------------------------------------------------------------------ planning
    def _plan(self, kind: str, *a):
        key = (kind,) + a
        if key not in self._plans:
            p = layout.window_plan(*a) if kind == "win" else layout.full_plan(*a)
            self._plans[key] = self.ops.make_plan(p)
        return self._plans[key]

    # ------------------------------------------------------------------ image encoder
    def set_image(self, pixel_values: torch.Tensor, capture: Optional[list] = None) -> None:
        """pixel_values fp32 [3, S, S] (normalised; S = cfg.image_size).  Leaves feat_s0 [N0, 32], feat_s1 [N1, 64] and
        image_embed [N2, 256] (+ no-memory and no-mask embeddings) on the device, all in row-major (y, x) token order."""
        if not (self.use_graph and capture is None and pixel_values.is_cuda):
            self._set_image(pixel_values, capture)
            self._pred_graphs.clear()       # prompt graphs read the features at their captured addresses
            self._img_graph = None
            return
        if self._img_graph is None or self._img_graph[1].shape != pixel_values.shape:
            static_in = pixel_values.detach().to(torch.float32).clone()
            self._set_image(static_in)      # warm-up outside capture: attention plans, function attributes, tensor maps
            self._pred_graphs.clear()
            g = self._capture(lambda: self._set_image(static_in))
            self._img_graph = (g, static_in, (self.feat_s0, self.feat_s1, self.image_embed))
        g, static_in, feats = self._img_graph
        static_in.copy_(pixel_values, non_blocking=True)
        g.replay()
        self.feat_s0, self.feat_s1, self.image_embed = feats

    @staticmethod
    def _capture(fn):
        """Capture ``fn`` (already warmed up) into a CUDA graph; kernel launches recorded by the capture are not counted."""
        from .. import lib
        g = torch.cuda.CUDAGraph()
        l0 = lib.launch_count
        with torch.cuda.graph(g):
            fn()
        lib.launch_count = l0
        return g

    def _set_image(self, pixel_values: torch.Tensor, capture: Optional[list] = None) -> None:
        cfg, W, ops = self.cfg, self.w, self.ops
        res = [cfg.grid >> s for s in range(4)]
        x = ops.patch_embed(pixel_values, W)                                   # [N0, C0], window-major(ws_0), + pos embed
        feats: List[torch.Tensor] = []
        stage_end = np.cumsum(cfg.blocks_per_stage) - 1
        for i, b in enumerate(W.blocks):
            s = b.stage
            r_in = res[s - 1] if b.pool else res[s]
            n_in = r_in * r_in
            h = ops.layernorm(x, b.ln1_w, b.ln1_b, cfg.layer_norm_eps)
            if b.pool:   # residual = max-pool(proj(LN1(x)))
                residual = ops.pool4(ops.linear(h, b.wres, b.bres), W.tables[f"pool{s}"])
            else:
                residual = x
            qk = ops.linear(h, b.wqk, b.bqk)                                   # [n_in, H*256]: q heads | k heads, 128-wide slots
            vt = ops.linear(b.wv, h)                                           # [H*hdv, n_in] = V^T (bias folded into bproj)
            H = b.heads
            q = qk[:, : H * 128]
            k = qk[:, H * 128:]
            n_q = n_in
            if b.pool:
                q = ops.pool4(q, W.tables[f"pool{s}"])                         # Q pooling: [n_in / 4, H*128]
                n_q = n_in // 4
            if b.window == 0:
                plan = self._plan("full", n_q, n_in)
            else:
                kv_per = b.window * b.window
                plan = self._plan("win", n_in // kv_per, kv_per // 4 if b.pool else kv_per, kv_per)
            ao = ops.attention(q.unflatten(1, (H, 128)), k.unflatten(1, (H, 128)), vt, plan, b.hd ** -0.5, b.hdv)   # [n_q, H*hdv]
            x = ops.linear(ao, b.wproj, b.bproj, residual=residual)
            h2 = ops.layernorm(x, b.ln2_w, b.ln2_b, cfg.layer_norm_eps)
            x = ops.linear(ops.linear(h2, b.w1, b.b1, act="gelu"), b.w2, b.b2, residual=x)
            if b.pool and f"reorder{s}" in W.tables:
                x = ops.gather_rows(x, W.tables[f"reorder{s}"])
            if capture is not None:
                capture.append((i, s, x))
            if i in stage_end:
                feats.append(ops.gather_rows(x, W.tables[f"to_rm{s}"]))        # stage output in row-major order
        # ---- neck: lateral 1x1 convs, top-down on the image-embedding level; conv_s0 / conv_s1 composed into the laterals
        lat3 = ops.linear(feats[3], W.lat3_w, W.lat3_b)
        lat2 = ops.linear(feats[2], W.lat2_w, W.lat2_b)
        self.image_embed = ops.upsample2x_add(lat2, lat3, res[3])              # [N2, 256]
        self.feat_s1 = ops.linear(feats[1], W.s1_w, W.s1_b)                    # [N1, 64]
        self.feat_s0 = ops.linear(feats[0], W.s0_w, W.s0_b)                    # [N0, 32]

    # ------------------------------------------------------------------ mask decoder
    def _attn(self, a: AttnW, q_in, k_in, v_in, heads: int):
        ops = self.ops
        q = ops.linear(q_in, a.wq, a.bq)
        k = ops.linear(k_in, a.wk, a.bk)
        v = ops.linear(v_in, a.wv, a.bv)
        return ops.small_attention(q, k, v, heads), a

    def predict_logits(self, coords: np.ndarray, labels: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
        """One object prompt.  coords fp32 [n, 2] in MODEL input pixels (x, y; already scaled to cfg.image_size), labels int [n]:
        1 / 0 foreground / background point, 2 / 3 box corners, -1 padding ("not a point").  Returns (low-res mask logits fp32
        [num_multimask, 4E, 4E], predicted IoU fp32 [num_multimask]) of the multi-mask outputs (tokens 1..3)."""
        assert self.image_embed is not None, "set_image first"
        if not (self.use_graph and self._img_graph is not None):
            return self._predict(coords, labels)
        coords = np.asarray(coords, np.float32).reshape(-1, 2)
        labels = np.asarray(labels, np.int32).reshape(-1)
        n = int(labels.shape[0])
        st = self._pred_graphs.get(n)
        dev = self.image_embed.device
        if st is None:
            c_host = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
            l_host = torch.empty((n,), dtype=torch.int32, pin_memory=True)
            c_host.copy_(torch.from_numpy(coords)); l_host.copy_(torch.from_numpy(labels))
            c_dev, l_dev = c_host.to(dev), l_host.to(dev)
            self._predict(c_dev, l_dev)       # warm-up outside capture
            out = []
            g = self._capture(lambda: out.extend(self._predict(c_dev, l_dev)))
            st = self._pred_graphs[n] = [g, c_host, l_host, c_dev, l_dev, out[0], out[1], None]
        g, c_host, l_host, c_dev, l_dev, logits, iou, uploaded = st
        if uploaded is not None:
            uploaded.synchronize()            # the previous prompt's upload has left the pinned pair
        c_host.copy_(torch.from_numpy(coords)); l_host.copy_(torch.from_numpy(labels))
        c_dev.copy_(c_host, non_blocking=True); l_dev.copy_(l_host, non_blocking=True)
        st[7] = torch.cuda.Event()
        st[7].record(torch.cuda.current_stream(dev))
        g.replay()
        return logits.clone(), iou.clone()

    def _predict(self, coords, labels) -> Tuple[torch.Tensor, torch.Tensor]:
        cfg, W, ops = self.cfg, self.w, self.ops
        Hh = cfg.dec_heads
        sparse = ops.prompt_embed(coords, labels, W, cfg.image_size)           # [n, 256]
        pe = ops.cat_rows(W.out_tokens, sparse)                                # query positional embedding = the initial tokens
        queries, keys, key_pe = pe, self.image_embed, W.image_pe
        eps = 1e-5   # nn.LayerNorm default inside the two-way transformer
        for li, L in enumerate(W.layers):
            if li == 0:   # skip_first_layer_pe: the self-attention output REPLACES the queries
                o, a = self._attn(L.self_attn, queries, queries, queries, Hh)
                queries = ops.linear(o, a.wo, a.bo)
            else:
                qp = ops.add(queries, pe)
                o, a = self._attn(L.self_attn, qp, qp, queries, Hh)
                queries = ops.linear(o, a.wo, a.bo, residual=queries)
            queries = ops.layernorm(queries, L.ln[0], L.ln[1], eps)
            # tokens -> image
            qp, kp = ops.add(queries, pe), ops.add(keys, key_pe)
            o, a = self._attn(L.t2i, qp, kp, keys, Hh)
            queries = ops.layernorm(ops.linear(o, a.wo, a.bo, residual=queries), L.ln[2], L.ln[3], eps)
            # MLP (ReLU)
            hid = ops.relu(ops.linear(queries, L.mlp_w1, L.mlp_b1))
            queries = ops.layernorm(ops.linear(hid, L.mlp_w2, L.mlp_b2, residual=queries), L.ln[4], L.ln[5], eps)
            # image -> tokens
            qp = ops.add(queries, pe)
            o, a = self._attn(L.i2t, kp, qp, queries, Hh)
            keys = ops.layernorm(ops.linear(o, a.wo, a.bo, residual=keys), L.ln[6], L.ln[7], eps)
        qp, kp = ops.add(queries, pe), ops.add(keys, key_pe)
        o, a = self._attn(W.final_attn, qp, kp, keys, Hh)
        queries = ops.layernorm(ops.linear(o, a.wo, a.bo, residual=queries), W.ln_final[0], W.ln_final[1], eps)
        # ---- upscaling: 2 x (transposed conv 2x2 stride 2 as a GEMM + pixel shuffle), high-resolution features added
        E = cfg.image_size // cfg.prompt_patch
        up = ops.upscale1(ops.linear(keys, W.up1_w), self.feat_s1, W.up1_b, W.up_ln_w, W.up_ln_b, E)          # [4 E^2, 64]
        up = ops.upscale2(ops.linear(up, W.up2_w), self.feat_s0, W.up2_b, 2 * E)                              # [16 E^2, 32]
        nm = cfg.num_multimask + 1
        hyper = []
        for t in range(nm):
            hw = W.hyper[t]
            v = queries[2 + t: 3 + t]
            for j in range(0, len(hw) - 2, 2):
                v = ops.relu(ops.linear(v, hw[j], hw[j + 1]))
            hyper.append(ops.linear(v, hw[-2], hw[-1]))
        hyper = ops.cat_rows(*hyper)                                           # [nm, 32]
        logits = ops.mask_logits(hyper, up, 1, cfg.num_multimask)              # fp32 [3, 16 E^2]
        v = queries[1:2]
        hw = W.iou_head
        for j in range(0, len(hw) - 2, 2):
            v = ops.relu(ops.linear(v, hw[j], hw[j + 1]))
        iou = ops.sigmoid(ops.linear(v, hw[-2], hw[-1], out_dtype=torch.float32))[0, 1:1 + cfg.num_multimask]
        return logits.view(cfg.num_multimask, 4 * E, 4 * E), iou
