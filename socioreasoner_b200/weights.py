"""Checkpoint layout -> kernel layout.

The checkpoint contract is the transformers-4.x key layout the reference's converter maps
(``mcore_adapter/src/mcore_adapter/models/converter/template.py`` ``register_template("qwen2_5_vl", ...)``;
SURVEY.md Appendix A): ``visual.*``, ``model.layers.*``, ``model.embed_tokens.weight``, ``model.norm.weight``,
``lm_head.weight``.  The 5.x in-memory names (``model.visual.*``, ``model.language_model.*``) are accepted too.

Kernel layout (all bf16 matrices are nn.Linear-style [N, K], K contiguous; norms / biases fp32):
  * patch embed Conv3d weight flattened to [1280, 1176] (columns (C,T,Ph,Pw) = pixel_values column order)
  * ViT MLP gate/up fused to one [2*Ipad, 1280] matrix, rows interleaved in groups of 16 gate / 16 up (SwiGLU is
    evaluated in the GEMM epilogue on 32-column TMEM chunks); I = 3420 zero-padded to Ipad = 3456 so every row is a
    multiple of 16 bytes (TMA) -- down_proj K is padded with zero columns to match
  * decoder q/k/v fused to [2560, 2048]; gate/up fused + interleaved to [22016, 2048]
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import torch

from .config import SocioReasonerConfig

GROUP = 16  # SwiGLU interleave granularity (rows)


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor, pad_to: int) -> torch.Tensor:
    """[I, ...] x2 -> [2*pad_to, ...] with rows [16 gate | 16 up] repeating; zero rows beyond I."""
    I = gate.shape[0]
    rest = gate.shape[1:]
    g = torch.zeros((pad_to,) + tuple(rest), dtype=gate.dtype, device=gate.device)
    u = torch.zeros_like(g)
    g[:I] = gate
    u[:I] = up
    g = g.view(pad_to // GROUP, 1, GROUP, *rest)
    u = u.view(pad_to // GROUP, 1, GROUP, *rest)
    return torch.cat([g, u], dim=1).reshape(2 * pad_to, *rest).contiguous()


def normalize_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        if k.startswith("model.visual."):
            k = "visual." + k[len("model.visual."):]
        elif k.startswith("model.language_model."):
            k = "model." + k[len("model.language_model."):]
        out[k] = v
    return out


@dataclass
class VitBlockW:
    norm1: torch.Tensor
    norm2: torch.Tensor
    qkv_w: torch.Tensor
    qkv_b: torch.Tensor
    proj_w: torch.Tensor
    proj_b: torch.Tensor
    gu_w: torch.Tensor
    gu_b: torch.Tensor
    down_w: torch.Tensor
    down_b: torch.Tensor


@dataclass
class DecLayerW:
    ln1: torch.Tensor
    ln2: torch.Tensor
    qkv_w: torch.Tensor
    qkv_b: torch.Tensor
    o_w: torch.Tensor
    gu_w: torch.Tensor
    down_w: torch.Tensor
    # second copy in the decode-streaming layout [N/16][K/32][16][32] (lib.pack_skinny); None = decode reads the [N,K] copy
    qkv_wp: Optional[torch.Tensor] = None
    o_wp: Optional[torch.Tensor] = None
    gu_wp: Optional[torch.Tensor] = None
    down_wp: Optional[torch.Tensor] = None
    # third copy of the down projection for the persistent decode kernel, sliced along K to match the gate/up units (see
    # ``_kslice_pack``): the CTA that produced 16 SwiGLU columns multiplies them by "its" 2048 x 16 slice right away
    down_wk: Optional[torch.Tensor] = None


@dataclass
class PackedWeights:
    patch_w: torch.Tensor = None
    vit_blocks: List[VitBlockW] = field(default_factory=list)
    merger_ln: torch.Tensor = None
    merger_w0: torch.Tensor = None
    merger_b0: torch.Tensor = None
    merger_w2: torch.Tensor = None
    merger_b2: torch.Tensor = None
    embed: torch.Tensor = None
    layers: List[DecLayerW] = field(default_factory=list)
    final_norm: torch.Tensor = None
    lm_head: torch.Tensor = None  # == embed when tied
    lm_head_p: Optional[torch.Tensor] = None  # decode-streaming copy

    def nbytes(self) -> int:
        seen, tot = set(), 0
        def add(t):
            nonlocal tot
            if t is not None and t.data_ptr() not in seen:
                seen.add(t.data_ptr()); tot += t.numel() * t.element_size()
        for t in (self.patch_w, self.merger_ln, self.merger_w0, self.merger_b0, self.merger_w2, self.merger_b2, self.embed,
                  self.final_norm, self.lm_head):
            add(t)
        for b in self.vit_blocks:
            for t in vars(b).values():
                add(t)
        for l in self.layers:
            for t in vars(l).values():
                add(t)
        return tot


def _stream_pack(w: torch.Tensor) -> torch.Tensor:
    N, K = w.shape
    return w.view(N // 16, 16, K // 32, 32).permute(0, 2, 1, 3).contiguous()


_KPERM = [2 * (p >> 2) + 8 * ((p >> 1) & 1) + (p & 1) for p in range(16)]   # position -> k inside a 16-wide k group


def _kslice_pack(w: torch.Tensor) -> torch.Tensor:
    """down_proj [2048, K] -> [K/16 units][4 stages][8 warps][4][16 rows][16 k]: unit u = the 16 input columns 16u.. (one gate/up
    unit of the persistent decode kernel), 64 KB contiguous, cut into 4 ring stages in which warp w finds the 16 x 16 blocks of its
    row groups 16w + 4s + j; inside a block every row's k values are stored in the order thread t of an mma.m16n8k16 quad reads
    them ({2t, 2t+1, 2t+8, 2t+9} contiguous), so that one 8-byte shared-memory load is half an A fragment."""
    N, K = w.shape
    assert N == 2048 and K % 16 == 0, "the persistent decode kernel fixes hidden = 2048"
    x = w.view(8, 4, 4, 16, K // 16, 16)[..., _KPERM]        # [w, s, j, r, u, pos]
    return x.permute(4, 1, 0, 2, 3, 5).contiguous()


def add_decode_streaming_copies(W: "PackedWeights") -> "PackedWeights":
    """Second copy of the decoder linears in the layout the weight-streaming decode kernel reads fastest (1 KB blocks of
    16 rows x 32 k: +6.2 GB for the 3B model, measured +12..20% HBM throughput at M = 8; 180 GB of HBM make this cheap)."""
    for L in W.layers:
        L.qkv_wp, L.o_wp, L.gu_wp, L.down_wp = _stream_pack(L.qkv_w), _stream_pack(L.o_w), _stream_pack(L.gu_w), _stream_pack(L.down_w)
        if L.down_w.shape[0] == 2048:
            L.down_wk = _kslice_pack(L.down_w)
    W.lm_head_p = _stream_pack(W.lm_head)
    return W


def pack(sd: Dict[str, torch.Tensor], cfg: SocioReasonerConfig, device="cuda", decode_streaming_copy: bool = True) -> PackedWeights:
    cfg.validate()
    sd = normalize_keys(sd)
    v, t = cfg.vision, cfg.text
    bf = lambda x: x.to(device=device, dtype=torch.bfloat16).contiguous()
    f32 = lambda x: x.to(device=device, dtype=torch.float32).contiguous()
    W = PackedWeights()
    W.patch_w = bf(sd["visual.patch_embed.proj.weight"].reshape(v.hidden_size, -1))
    Ip = v.intermediate_padded
    for i in range(v.depth):
        p = f"visual.blocks.{i}."
        gate_w, up_w = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
        gate_b, up_b = sd[p + "mlp.gate_proj.bias"], sd[p + "mlp.up_proj.bias"]
        down = sd[p + "mlp.down_proj.weight"]
        down_p = torch.zeros((down.shape[0], Ip), dtype=down.dtype, device=down.device)
        down_p[:, : v.intermediate_size] = down
        W.vit_blocks.append(VitBlockW(
            norm1=f32(sd[p + "norm1.weight"]), norm2=f32(sd[p + "norm2.weight"]),
            qkv_w=bf(sd[p + "attn.qkv.weight"]), qkv_b=f32(sd[p + "attn.qkv.bias"]),
            proj_w=bf(sd[p + "attn.proj.weight"]), proj_b=f32(sd[p + "attn.proj.bias"]),
            gu_w=bf(interleave_gate_up(gate_w, up_w, Ip)), gu_b=f32(interleave_gate_up(gate_b, up_b, Ip)),
            down_w=bf(down_p), down_b=f32(sd[p + "mlp.down_proj.bias"]),
        ))
    W.merger_ln = f32(sd["visual.merger.ln_q.weight"])
    W.merger_w0, W.merger_b0 = bf(sd["visual.merger.mlp.0.weight"]), f32(sd["visual.merger.mlp.0.bias"])
    W.merger_w2, W.merger_b2 = bf(sd["visual.merger.mlp.2.weight"]), f32(sd["visual.merger.mlp.2.bias"])
    W.embed = bf(sd["model.embed_tokens.weight"])
    for i in range(t.num_hidden_layers):
        p = f"model.layers.{i}."
        qkv_w = torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.v_proj.weight"]], 0)
        qkv_b = torch.cat([sd[p + "self_attn.q_proj.bias"], sd[p + "self_attn.k_proj.bias"], sd[p + "self_attn.v_proj.bias"]], 0)
        W.layers.append(DecLayerW(
            ln1=f32(sd[p + "input_layernorm.weight"]), ln2=f32(sd[p + "post_attention_layernorm.weight"]),
            qkv_w=bf(qkv_w), qkv_b=f32(qkv_b), o_w=bf(sd[p + "self_attn.o_proj.weight"]),
            gu_w=bf(interleave_gate_up(sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"], t.intermediate_size)),
            down_w=bf(sd[p + "mlp.down_proj.weight"]),
        ))
    W.final_norm = f32(sd["model.norm.weight"])
    if t.tie_word_embeddings or "lm_head.weight" not in sd:
        W.lm_head = W.embed
    else:
        lh = sd["lm_head.weight"]
        W.lm_head = W.embed if lh.data_ptr() == sd["model.embed_tokens.weight"].data_ptr() else bf(lh)
    if decode_streaming_copy:
        add_decode_streaming_copies(W)
    return W


def random_state_dict(cfg: SocioReasonerConfig, seed: int = 42, device="cuda", std: float = 0.02,
                      dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Random-init weights of the architecture in checkpoint (4.x) layout, generated on ``device``
    (N(0, 0.02) matrices; norm weights ~ 1, small random biases so every code path carries signal)."""
    g = torch.Generator(device=device).manual_seed(seed)
    v, t = cfg.vision, cfg.text
    rn = lambda *s: (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * std).to(dtype)
    nw = lambda n: (1.0 + 0.1 * torch.randn(n, generator=g, device=device, dtype=torch.float32)).to(dtype)
    sd: Dict[str, torch.Tensor] = {}
    sd["visual.patch_embed.proj.weight"] = rn(v.hidden_size, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size)
    for i in range(v.depth):
        p = f"visual.blocks.{i}."
        sd[p + "norm1.weight"], sd[p + "norm2.weight"] = nw(v.hidden_size), nw(v.hidden_size)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = rn(3 * v.hidden_size, v.hidden_size), rn(3 * v.hidden_size)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = rn(v.hidden_size, v.hidden_size), rn(v.hidden_size)
        for n in ("gate_proj", "up_proj"):
            sd[p + f"mlp.{n}.weight"], sd[p + f"mlp.{n}.bias"] = rn(v.intermediate_size, v.hidden_size), rn(v.intermediate_size)
        sd[p + "mlp.down_proj.weight"], sd[p + "mlp.down_proj.bias"] = rn(v.hidden_size, v.intermediate_size), rn(v.hidden_size)
    hm = v.hidden_size * v.spatial_merge_size ** 2
    sd["visual.merger.ln_q.weight"] = nw(v.hidden_size)
    sd["visual.merger.mlp.0.weight"], sd["visual.merger.mlp.0.bias"] = rn(hm, hm), rn(hm)
    sd["visual.merger.mlp.2.weight"], sd["visual.merger.mlp.2.bias"] = rn(v.out_hidden_size, hm), rn(v.out_hidden_size)
    sd["model.embed_tokens.weight"] = rn(t.vocab_size, t.hidden_size)
    kvd = t.num_key_value_heads * t.head_dim
    for i in range(t.num_hidden_layers):
        p = f"model.layers.{i}."
        sd[p + "input_layernorm.weight"], sd[p + "post_attention_layernorm.weight"] = nw(t.hidden_size), nw(t.hidden_size)
        sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"] = rn(t.hidden_size, t.hidden_size), rn(t.hidden_size)
        sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"] = rn(kvd, t.hidden_size), rn(kvd)
        sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"] = rn(kvd, t.hidden_size), rn(kvd)
        sd[p + "self_attn.o_proj.weight"] = rn(t.hidden_size, t.hidden_size)
        sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"] = rn(t.intermediate_size, t.hidden_size), rn(t.intermediate_size, t.hidden_size)
        sd[p + "mlp.down_proj.weight"] = rn(t.hidden_size, t.intermediate_size)
    sd["model.norm.weight"] = nw(t.hidden_size)
    return sd


# ---------------------------------------------------------------------------------------------------------------------
# In-place weight update (train -> infer weight sync).  The reference pushes the trained parameters into the rollout engine
# tensor by tensor under their HF checkpoint names (``worker_helper.py:104-116`` -> ``load_weights([(name, weight)])``) every
# pipeline step (``rlvr_socioseg_vlm_pipeline_infer.py:608-612,634``).  Here each name maps to a slice of a kernel-layout
# tensor; every destination is overwritten IN PLACE, so device pointers held by the decode layer table and by captured CUDA
# graphs stay valid.
def _put_interleaved(dst: torch.Tensor, src: torch.Tensor, which: int) -> None:
    """rows of ``src`` [I, ...] -> the gate (which=0) or up (which=1) rows of ``dst`` [2*Ipad, ...] ([16 gate | 16 up] groups)."""
    I = src.shape[0]
    rest = tuple(dst.shape[1:])
    view = dst.view(dst.shape[0] // (2 * GROUP), 2, GROUP, *rest)
    full = I // GROUP
    src = src.to(device=dst.device, dtype=dst.dtype)
    if full:
        view[:full, which].copy_(src[: full * GROUP].reshape(full, GROUP, *rest))
    if I % GROUP:
        view[full, which, : I % GROUP].copy_(src[full * GROUP:])


def update_weights(W: PackedWeights, named_tensors, cfg: SocioReasonerConfig) -> List[str]:
    """Overwrite the kernel-layout weights from ``(hf_name, tensor)`` pairs (any subset, any order).  Returns the names
    consumed; an unknown name raises ``KeyError`` (the reference's ``load_weights`` fails the same way on a foreign key)."""
    v, t = cfg.vision, cfg.text
    H, kvd = t.hidden_size, t.num_key_value_heads * t.head_dim
    done: List[str] = []
    dirty = set()  # (layer index, matrix) whose decode-streaming copy must be re-packed
    with torch.no_grad():
        for raw_name, w in named_tensors:
            name = next(iter(normalize_keys({raw_name: None})))
            put = lambda dst, src: dst.copy_(src.to(device=dst.device, dtype=dst.dtype).reshape(dst.shape))
            parts = name.split(".")
            if name == "visual.patch_embed.proj.weight":
                put(W.patch_w, w)
            elif name.startswith("visual.blocks."):
                B, leaf = W.vit_blocks[int(parts[2])], ".".join(parts[3:])
                simple = {"norm1.weight": B.norm1, "norm2.weight": B.norm2, "attn.qkv.weight": B.qkv_w, "attn.qkv.bias": B.qkv_b,
                          "attn.proj.weight": B.proj_w, "attn.proj.bias": B.proj_b, "mlp.down_proj.bias": B.down_b}
                if leaf in simple:
                    put(simple[leaf], w)
                elif leaf in ("mlp.gate_proj.weight", "mlp.up_proj.weight"):
                    _put_interleaved(B.gu_w, w, 0 if "gate" in leaf else 1)
                elif leaf in ("mlp.gate_proj.bias", "mlp.up_proj.bias"):
                    _put_interleaved(B.gu_b, w, 0 if "gate" in leaf else 1)
                elif leaf == "mlp.down_proj.weight":
                    B.down_w[:, : v.intermediate_size].copy_(w.to(device=B.down_w.device, dtype=B.down_w.dtype))
                else:
                    raise KeyError(raw_name)
            elif name.startswith("visual.merger."):
                leaf = ".".join(parts[2:])
                put({"ln_q.weight": W.merger_ln, "mlp.0.weight": W.merger_w0, "mlp.0.bias": W.merger_b0, "mlp.2.weight": W.merger_w2,
                     "mlp.2.bias": W.merger_b2}[leaf], w)
            elif name == "model.embed_tokens.weight":
                put(W.embed, w)
                if W.lm_head is W.embed:
                    dirty.add((-1, "lm_head"))
            elif name == "lm_head.weight":
                if W.lm_head is not W.embed:  # tied head: the embedding update already covers it
                    put(W.lm_head, w)
                    dirty.add((-1, "lm_head"))
            elif name == "model.norm.weight":
                put(W.final_norm, w)
            elif name.startswith("model.layers."):
                li = int(parts[2])
                L, leaf = W.layers[li], ".".join(parts[3:])
                rows = {"q_proj": (0, H), "k_proj": (H, H + kvd), "v_proj": (H + kvd, H + 2 * kvd)}
                if leaf in ("input_layernorm.weight", "post_attention_layernorm.weight"):
                    put(L.ln1 if leaf.startswith("input") else L.ln2, w)
                elif parts[3] == "self_attn" and parts[4] in rows:
                    a, b = rows[parts[4]]
                    if parts[5] == "weight":
                        put(L.qkv_w[a:b], w)
                        dirty.add((li, "qkv"))
                    else:
                        put(L.qkv_b[a:b], w)
                elif leaf == "self_attn.o_proj.weight":
                    put(L.o_w, w)
                    dirty.add((li, "o"))
                elif leaf in ("mlp.gate_proj.weight", "mlp.up_proj.weight"):
                    _put_interleaved(L.gu_w, w, 0 if "gate" in leaf else 1)
                    dirty.add((li, "gu"))
                elif leaf == "mlp.down_proj.weight":
                    put(L.down_w, w)
                    dirty.add((li, "down"))
                else:
                    raise KeyError(raw_name)
            else:
                raise KeyError(raw_name)
            done.append(raw_name)
        for li, which in sorted(dirty):
            if li < 0:
                if W.lm_head_p is not None:
                    W.lm_head_p.copy_(_stream_pack(W.lm_head))
                continue
            L = W.layers[li]
            src, dst = {"qkv": (L.qkv_w, L.qkv_wp), "o": (L.o_w, L.o_wp), "gu": (L.gu_w, L.gu_wp), "down": (L.down_w, L.down_wp)}[which]
            if dst is not None:
                dst.copy_(_stream_pack(src))
            if which == "down" and L.down_wk is not None:
                L.down_wk.copy_(_kslice_pack(L.down_w))
    return done
