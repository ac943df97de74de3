"""SAM2 weights in kernel layout, built once from an HF-named state dict (``transformers.models.sam2.Sam2Model``: the names a
converted ``sam2_hiera_large`` checkpoint carries).  Pure re-layout plus a few exact algebraic folds, all done at load time:

  * q / k projection rows are spread to one 128-wide slot per head (head_dim 72 .. 112 zero-padded) so that the GEMM writes the
    attention kernel's operand layout directly; the v projection is kept as the LEFT operand of a second GEMM (V^T = Wv . X^T),
    which produces the K-major V^T the P.V MMA wants without a transpose pass;
  * the v bias leaves the attention (softmax rows sum to one: P.(V + 1 b^T) = P.V + b) and is folded into the output projection's
    bias (b_proj + W_proj b_v); the projection's input columns are padded per head like V;
  * the neck's lateral 1x1 convolutions of the two high-resolution levels are composed with the decoder's conv_s0 / conv_s1 (no
    non-linearity in between); ``no_memory_embedding`` and the prompt encoder's ``no_mask_embed`` (dense prompt when no mask
    is given) become part of the image-embedding level's bias;
  * the windowed positional embedding (bicubic-interpolated background + tiled window table, HF ``_get_pos_embed``) is tabulated
    in the trunk's window-major token order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import layout
from .config import Sam2Cfg


def _hdv(hd: int) -> int:
    return 80 if hd <= 80 else 128


@dataclass
class HieraBlock:
    stage: int
    dim_in: int
    dim: int
    heads: int
    hd: int
    hdv: int
    pool: bool            # first block of a stage > 0: q-pooling + projected, pooled residual
    window: int           # 0 = global attention; else the window size in the resolution of this block's INPUT
    ln1_w: torch.Tensor = None
    ln1_b: torch.Tensor = None
    wqk: torch.Tensor = None      # [heads*256, dim_in]
    bqk: torch.Tensor = None
    wv: torch.Tensor = None       # [heads*hdv, dim_in]
    wproj: torch.Tensor = None    # [dim, heads*hdv]
    bproj: torch.Tensor = None    # proj bias + W_proj b_v
    wres: Optional[torch.Tensor] = None
    bres: Optional[torch.Tensor] = None
    ln2_w: torch.Tensor = None
    ln2_b: torch.Tensor = None
    w1: torch.Tensor = None
    b1: torch.Tensor = None
    w2: torch.Tensor = None
    b2: torch.Tensor = None


@dataclass
class AttnW:
    wq: torch.Tensor
    bq: torch.Tensor
    wk: torch.Tensor
    bk: torch.Tensor
    wv: torch.Tensor
    bv: torch.Tensor
    wo: torch.Tensor
    bo: torch.Tensor


@dataclass
class DecLayer:
    self_attn: AttnW
    t2i: AttnW
    i2t: AttnW
    ln: List[torch.Tensor]        # [w1,b1,w2,b2,w3,b3,w4,b4]
    mlp_w1: torch.Tensor
    mlp_b1: torch.Tensor
    mlp_w2: torch.Tensor
    mlp_b2: torch.Tensor


@dataclass
class Sam2Weights:
    patch_w: torch.Tensor = None          # fp32 [C0, 3*k*k]
    patch_b: torch.Tensor = None
    pos_table: torch.Tensor = None        # fp32 [N0, C0], window-major order of stage 0
    patch_rows: torch.Tensor = None       # int32 [N0]: destination row of row-major pixel p
    blocks: List[HieraBlock] = field(default_factory=list)
    lat3_w: torch.Tensor = None
    lat3_b: torch.Tensor = None
    lat2_w: torch.Tensor = None
    lat2_b: torch.Tensor = None           # + no_memory_embedding + no_mask_embed
    s1_w: torch.Tensor = None             # conv_s1 o neck lateral (stage 2 features)
    s1_b: torch.Tensor = None
    s0_w: torch.Tensor = None
    s0_b: torch.Tensor = None
    image_pe: torch.Tensor = None         # bf16 [N2, 256]
    # prompt encoder
    pe_gauss: torch.Tensor = None         # fp32 [2, 128]
    point_embed: torch.Tensor = None      # fp32 [4, 256]
    not_a_point: torch.Tensor = None      # fp32 [256]
    # decoder
    out_tokens: torch.Tensor = None       # bf16 [2 + num_mask_tokens, 256]: obj score, iou, mask tokens
    layers: List[DecLayer] = field(default_factory=list)
    final_attn: AttnW = None
    ln_final: List[torch.Tensor] = None
    up1_w: torch.Tensor = None            # [4*64, 256]  rows (dy*2+dx)*64 + co
    up1_b: torch.Tensor = None            # fp32 [64]
    up_ln_w: torch.Tensor = None
    up_ln_b: torch.Tensor = None
    up2_w: torch.Tensor = None            # [4*32, 64]
    up2_b: torch.Tensor = None
    hyper: List[List[torch.Tensor]] = field(default_factory=list)   # per mask token: [w1,b1,w2,b2,w3,b3]
    iou_head: List[torch.Tensor] = None
    # index tables
    tables: Dict[str, torch.Tensor] = field(default_factory=dict)


def block_specs(cfg: Sam2Cfg) -> List[HieraBlock]:
    out, total = [], 0
    for s, nb in enumerate(cfg.blocks_per_stage):
        for b in range(nb):
            first = s > 0 and b == 0
            dim_in = cfg.embed_dims[s - 1] if first else cfg.embed_dims[s]
            window = cfg.window_sizes[s - 1] if first else cfg.window_sizes[s]
            if total in cfg.global_attention_blocks:
                window = 0
            hd = cfg.embed_dims[s] // cfg.num_heads[s]
            out.append(HieraBlock(stage=s, dim_in=dim_in, dim=cfg.embed_dims[s], heads=cfg.num_heads[s], hd=hd, hdv=_hdv(hd), pool=first,
                                  window=window))
            total += 1
    return out


def pack(sd: Dict[str, torch.Tensor], cfg: Sam2Cfg, device, act_dtype=torch.bfloat16) -> Sam2Weights:
    cfg.validate()
    dev = torch.device(device)
    f32 = lambda t: t.detach().to(torch.float32)
    W = lambda t: f32(t).to(act_dtype).contiguous().to(dev)       # matrices: activation dtype
    V = lambda t: f32(t).contiguous().to(dev)                       # biases / norm parameters / tables: fp32
    out = Sam2Weights()
    p = "vision_encoder.backbone."
    C0, R0 = cfg.embed_dims[0], cfg.grid
    out.patch_w = V(sd[p + "patch_embed.projection.weight"].reshape(C0, -1))
    out.patch_b = V(sd[p + "patch_embed.projection.bias"])
    pos = F.interpolate(f32(sd[p + "pos_embed"]), size=(R0, R0), mode="bicubic")
    win = f32(sd[p + "pos_embed_window"])
    pos = pos + win.tile([x // y for x, y in zip(pos.shape, win.shape)])
    perm0 = layout.wm_perm(R0, cfg.window_sizes[0])
    out.pos_table = V(pos[0].permute(1, 2, 0).reshape(R0 * R0, C0)[torch.from_numpy(perm0).long()])
    out.patch_rows = torch.from_numpy(layout.inverse(perm0)).to(dev)
    # ---- trunk
    for i, b in enumerate(block_specs(cfg)):
        q = f"{p}blocks.{i}."
        C, H, hd, hdv = b.dim, b.heads, b.hd, b.hdv
        wqkv, bqkv = f32(sd[q + "attn.qkv.weight"]), f32(sd[q + "attn.qkv.bias"])
        wqk = torch.zeros(H * 256, b.dim_in)
        bqk = torch.zeros(H * 256)
        wv = torch.zeros(H * hdv, b.dim_in)
        wproj = torch.zeros(C, H * hdv)
        wp = f32(sd[q + "attn.proj.weight"])
        for h in range(H):
            wqk[h * 128:h * 128 + hd] = wqkv[h * hd:(h + 1) * hd]
            bqk[h * 128:h * 128 + hd] = bqkv[h * hd:(h + 1) * hd]
            wqk[H * 128 + h * 128:H * 128 + h * 128 + hd] = wqkv[C + h * hd:C + (h + 1) * hd]
            bqk[H * 128 + h * 128:H * 128 + h * 128 + hd] = bqkv[C + h * hd:C + (h + 1) * hd]
            wv[h * hdv:h * hdv + hd] = wqkv[2 * C + h * hd:2 * C + (h + 1) * hd]
            wproj[:, h * hdv:h * hdv + hd] = wp[:, h * hd:(h + 1) * hd]
        b.ln1_w, b.ln1_b = V(sd[q + "layer_norm1.weight"]), V(sd[q + "layer_norm1.bias"])
        b.ln2_w, b.ln2_b = V(sd[q + "layer_norm2.weight"]), V(sd[q + "layer_norm2.bias"])
        b.wqk, b.bqk, b.wv, b.wproj = W(wqk), V(bqk), W(wv), W(wproj)
        b.bproj = V(f32(sd[q + "attn.proj.bias"]) + wp @ bqkv[2 * C:])
        if b.pool:
            b.wres, b.bres = W(sd[q + "proj.weight"]), V(sd[q + "proj.bias"])
        b.w1, b.b1 = W(sd[q + "mlp.proj_in.weight"]), V(sd[q + "mlp.proj_in.bias"])
        b.w2, b.b2 = W(sd[q + "mlp.proj_out.weight"]), V(sd[q + "mlp.proj_out.bias"])
        out.blocks.append(b)
    # ---- neck (+ decoder convs on the high-resolution levels)
    n = "vision_encoder.neck.convs."
    conv = lambda k: (f32(sd[k + ".weight"]).flatten(1), f32(sd[k + ".bias"]))
    w3, b3 = conv(n + "0")
    w2, b2 = conv(n + "1")
    w1, b1 = conv(n + "2")
    w0, b0 = conv(n + "3")
    ws1, bs1 = conv("mask_decoder.conv_s1")
    ws0, bs0 = conv("mask_decoder.conv_s0")
    out.lat3_w, out.lat3_b = W(w3), V(b3)
    out.lat2_w = W(w2)
    out.lat2_b = V(b2 + f32(sd["no_memory_embedding"]).reshape(-1) + f32(sd["prompt_encoder.no_mask_embed.weight"]).reshape(-1))
    out.s1_w, out.s1_b = W(ws1 @ w1), V(ws1 @ b1 + bs1)
    out.s0_w, out.s0_b = W(ws0 @ w0), V(ws0 @ b0 + bs0)
    # ---- prompt encoder + image-wide positional encoding (HF get_image_wide_positional_embeddings)
    g = f32(sd["prompt_encoder.shared_embedding.positional_embedding"])
    out.pe_gauss = V(g)
    out.point_embed = V(sd["prompt_encoder.point_embed.weight"])
    out.not_a_point = V(sd["prompt_encoder.not_a_point_embed.weight"].reshape(-1))
    E = cfg.image_size // cfg.prompt_patch
    gi = f32(sd["shared_image_embedding.positional_embedding"])
    grid = torch.ones(E, E)
    y = (grid.cumsum(0) - 0.5) / E
    x = (grid.cumsum(1) - 0.5) / E
    c = (2 * torch.stack([x, y], -1) - 1) @ gi
    c = 2 * np.pi * c
    out.image_pe = W(torch.cat([c.sin(), c.cos()], -1).reshape(E * E, -1))
    # ---- decoder
    d = "mask_decoder."
    out.out_tokens = W(torch.cat([f32(sd[d + "obj_score_token.weight"]), f32(sd[d + "iou_token.weight"]), f32(sd[d + "mask_tokens.weight"])], 0))

    def attn(k):
        return AttnW(*[f(sd[f"{k}.{nm}_proj.{wb}"]) for nm in ("q", "k", "v", "o") for f, wb in ((W, "weight"), (V, "bias"))])

    for l in range(cfg.dec_layers):
        k = f"{d}transformer.layers.{l}."
        ln = [V(sd[f"{k}layer_norm{j}.{wb}"]) for j in (1, 2, 3, 4) for wb in ("weight", "bias")]
        out.layers.append(DecLayer(attn(k + "self_attn"), attn(k + "cross_attn_token_to_image"), attn(k + "cross_attn_image_to_token"), ln,
                                   W(sd[k + "mlp.proj_in.weight"]), V(sd[k + "mlp.proj_in.bias"]), W(sd[k + "mlp.proj_out.weight"]),
                                   V(sd[k + "mlp.proj_out.bias"])))
    out.final_attn = attn(d + "transformer.final_attn_token_to_image")
    out.ln_final = [V(sd[d + "transformer.layer_norm_final_attn.weight"]), V(sd[d + "transformer.layer_norm_final_attn.bias"])]
    u1 = f32(sd[d + "upscale_conv1.weight"])          # [in, out, 2, 2]
    u2 = f32(sd[d + "upscale_conv2.weight"])
    out.up1_w = W(u1.permute(2, 3, 1, 0).reshape(4 * u1.shape[1], u1.shape[0]))
    out.up2_w = W(u2.permute(2, 3, 1, 0).reshape(4 * u2.shape[1], u2.shape[0]))
    out.up1_b, out.up2_b = V(sd[d + "upscale_conv1.bias"]), V(sd[d + "upscale_conv2.bias"])
    out.up_ln_w, out.up_ln_b = V(sd[d + "upscale_layer_norm.weight"]), V(sd[d + "upscale_layer_norm.bias"])

    def mlp3(k):
        names = ["proj_in"] + [f"layers.{j}" for j in range(cfg.iou_depth - 2)] + ["proj_out"]
        ts = [f32(sd[f"{k}.{nm}.{wb}"]) for nm in names for wb in ("weight", "bias")]
        pad = (-ts[-2].shape[0]) % 8        # the GEMM wants N % 8 == 0: zero rows behind the last layer (IoU head: 4 outputs)
        if pad:
            ts[-2] = torch.cat([ts[-2], torch.zeros(pad, ts[-2].shape[1])], 0)
            ts[-1] = torch.cat([ts[-1], torch.zeros(pad)], 0)
        return [(W if j % 2 == 0 else V)(t) for j, t in enumerate(ts)]

    for t in range(cfg.num_multimask + 1):
        out.hyper.append(mlp3(f"{d}output_hypernetworks_mlps.{t}"))
    out.iou_head = mlp3(d + "iou_prediction_head")
    # ---- index tables
    R = [cfg.grid >> s for s in range(4)]
    ws = cfg.window_sizes
    T = out.tables
    for s in range(1, 4):
        T[f"pool{s}"] = torch.from_numpy(layout.pool_sources(R[s - 1], ws[s - 1])).to(dev)
        if ws[s] != ws[s - 1] // 2:   # re-order from the pooled order (windows of ws_prev / 2) to this stage's own windows
            src = layout.inverse(layout.wm_perm(R[s], ws[s - 1] // 2))[layout.wm_perm(R[s], ws[s])]
            T[f"reorder{s}"] = torch.from_numpy(src.astype(np.int32)).to(dev)
    for s in range(4):   # stage outputs back to row-major (for the neck)
        T[f"to_rm{s}"] = torch.from_numpy(layout.inverse(layout.wm_perm(R[s], ws[s]))).to(dev)
    return out


def random_state_dict(cfg: Sam2Cfg, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """HF-named random weights of the right shapes (bench / smoke without a checkpoint: there is no network on the box)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, n, k, bias=True):
        sd[name + ".weight"] = torch.randn(n, k, generator=g) * std
        if bias:
            sd[name + ".bias"] = torch.randn(n, generator=g) * std

    def norm(name, n):
        sd[name + ".weight"] = 1 + torch.randn(n, generator=g) * 0.1
        sd[name + ".bias"] = torch.randn(n, generator=g) * 0.05

    p = "vision_encoder.backbone."
    C0, k = cfg.embed_dims[0], cfg.patch_kernel
    sd[p + "patch_embed.projection.weight"] = torch.randn(C0, 3, k, k, generator=g) * std
    sd[p + "patch_embed.projection.bias"] = torch.randn(C0, generator=g) * std
    sd[p + "pos_embed"] = torch.randn(1, C0, *cfg.pos_embed_bkg, generator=g) * std
    sd[p + "pos_embed_window"] = torch.randn(1, C0, cfg.window_sizes[0], cfg.window_sizes[0], generator=g) * std
    for i, b in enumerate(block_specs(cfg)):
        q = f"{p}blocks.{i}."
        norm(q + "layer_norm1", b.dim_in)
        lin(q + "attn.qkv", 3 * b.dim, b.dim_in)
        lin(q + "attn.proj", b.dim, b.dim)
        norm(q + "layer_norm2", b.dim)
        lin(q + "mlp.proj_in", int(b.dim * cfg.mlp_ratio), b.dim)
        lin(q + "mlp.proj_out", b.dim, int(b.dim * cfg.mlp_ratio))
        if b.pool:
            lin(q + "proj", b.dim, b.dim_in)
    for j, c in enumerate(reversed(cfg.embed_dims)):
        sd[f"vision_encoder.neck.convs.{j}.weight"] = torch.randn(cfg.fpn_hidden, c, 1, 1, generator=g) * std
        sd[f"vision_encoder.neck.convs.{j}.bias"] = torch.randn(cfg.fpn_hidden, generator=g) * std
    Hd = cfg.dec_hidden
    sd["no_memory_embedding"] = torch.randn(1, 1, Hd, generator=g) * std
    sd["shared_image_embedding.positional_embedding"] = torch.randn(2, Hd // 2, generator=g)
    sd["prompt_encoder.shared_embedding.positional_embedding"] = torch.randn(2, Hd // 2, generator=g)
    sd["prompt_encoder.no_mask_embed.weight"] = torch.randn(1, Hd, generator=g) * std
    sd["prompt_encoder.point_embed.weight"] = torch.randn(4, Hd, generator=g) * std
    sd["prompt_encoder.not_a_point_embed.weight"] = torch.randn(1, Hd, generator=g) * std
    d = "mask_decoder."
    sd[d + "iou_token.weight"] = torch.randn(1, Hd, generator=g) * std
    sd[d + "obj_score_token.weight"] = torch.randn(1, Hd, generator=g) * std
    sd[d + "mask_tokens.weight"] = torch.randn(cfg.num_multimask + 1, Hd, generator=g) * std

    def attn(name, internal):
        for nm in ("q", "k", "v"):
            lin(f"{name}.{nm}_proj", internal, Hd)
        lin(f"{name}.o_proj", Hd, internal)

    for l in range(cfg.dec_layers):
        kk = f"{d}transformer.layers.{l}."
        attn(kk + "self_attn", Hd)
        attn(kk + "cross_attn_token_to_image", Hd // cfg.dec_downsample)
        attn(kk + "cross_attn_image_to_token", Hd // cfg.dec_downsample)
        for j in (1, 2, 3, 4):
            norm(f"{kk}layer_norm{j}", Hd)
        lin(kk + "mlp.proj_in", cfg.dec_mlp, Hd)
        lin(kk + "mlp.proj_out", Hd, cfg.dec_mlp)
    attn(d + "transformer.final_attn_token_to_image", Hd // cfg.dec_downsample)
    norm(d + "transformer.layer_norm_final_attn", Hd)
    sd[d + "upscale_conv1.weight"] = torch.randn(Hd, Hd // 4, 2, 2, generator=g) * std
    sd[d + "upscale_conv1.bias"] = torch.randn(Hd // 4, generator=g) * std
    sd[d + "upscale_conv2.weight"] = torch.randn(Hd // 4, Hd // 8, 2, 2, generator=g) * std
    sd[d + "upscale_conv2.bias"] = torch.randn(Hd // 8, generator=g) * std
    norm(d + "upscale_layer_norm", Hd // 4)

    def mlp(name, hidden, n_out, depth):
        lin(name + ".proj_in", hidden, Hd)
        for j in range(depth - 2):
            lin(f"{name}.layers.{j}", hidden, hidden)
        lin(name + ".proj_out", n_out, hidden)

    for t in range(cfg.num_multimask + 1):
        mlp(f"{d}output_hypernetworks_mlps.{t}", Hd, Hd // 8, cfg.iou_depth)
    mlp(d + "iou_prediction_head", cfg.iou_hidden, cfg.num_multimask + 1, cfg.iou_depth)
    lin(d + "conv_s0", Hd // 8, Hd)
    lin(d + "conv_s1", Hd // 4, Hd)
    sd[d + "conv_s0.weight"] = sd[d + "conv_s0.weight"].reshape(Hd // 8, Hd, 1, 1)
    sd[d + "conv_s1.weight"] = sd[d + "conv_s1.weight"].reshape(Hd // 4, Hd, 1, 1)
    return sd
