"""HF-eager oracle of the SocioReasoner VLM forward (TEST INFRASTRUCTURE, never on the product path).

Drives ``transformers``' Qwen2.5-VL exactly like the reference's eager strategy does
(``roll/distributed/strategy/hf_strategy.py:60-91``: explicit ``position_ids`` transposed to [3,B,S],
``use_cache=False``, concatenated ``pixel_values`` / ``image_grid_thw``) with seeded random-init
weights.  API drift of transformers 5.5 vs. the reference's 4.4x era is
neutralised here: ``pooler_output`` for merged ViT embeddings; the in-tree
``get_rope_index`` (``oracle/host_ref.py``) instead of HF's; nested ``text_config``.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import host_ref

# Qwen2.5-VL-3B dimensions.  `tiny*` configs keep every per-layer dimension
# (so every kernel sees the production K/N sizes) and only cut depth / vocab so CPU runs take seconds.
_TEXT_3B = dict(
    hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16,
    num_key_value_heads=2, vocab_size=151936, rms_norm_eps=1e-6, max_position_embeddings=128000,
    tie_word_embeddings=True,
    rope_parameters=dict(rope_type="default", rope_theta=1000000.0, mrope_section=[16, 24, 24]),
)
_VISION_3B = dict(
    depth=32, hidden_size=1280, intermediate_size=3420, num_heads=16, patch_size=14, spatial_merge_size=2,
    temporal_patch_size=2, window_size=112, out_hidden_size=2048, fullatt_block_indexes=[7, 15, 23, 31],
    hidden_act="silu", tokens_per_second=2,
)
_TOKENS_3B = dict(image_token_id=151655, video_token_id=151656, vision_start_token_id=151652,
                  vision_end_token_id=151653)

CONFIGS: Dict[str, dict] = {
    "3b": dict(text=_TEXT_3B, vision=_VISION_3B, tokens=_TOKENS_3B),
    # 2 ViT blocks (block 0 windowed, block 1 full attention), 2 decoder layers, 8k vocab.
    "tiny": dict(
        text={**_TEXT_3B, "num_hidden_layers": 2, "vocab_size": 8192},
        vision={**_VISION_3B, "depth": 2, "fullatt_block_indexes": [1]},
        tokens=dict(image_token_id=8001, video_token_id=8002, vision_start_token_id=8003, vision_end_token_id=8004),
    ),
    # 4 ViT blocks / 4 decoder layers: catches layer-to-layer drift that 2 layers cannot.
    "small": dict(
        text={**_TEXT_3B, "num_hidden_layers": 4, "vocab_size": 8192},
        vision={**_VISION_3B, "depth": 4, "fullatt_block_indexes": [1, 3]},
        tokens=dict(image_token_id=8001, video_token_id=8002, vision_start_token_id=8003, vision_end_token_id=8004),
    ),
}


def make_config(name: str):
    from transformers import Qwen2_5_VLConfig

    c = CONFIGS[name]
    return Qwen2_5_VLConfig(
        text_config=copy.deepcopy(c["text"]), vision_config=copy.deepcopy(c["vision"]),
        tie_word_embeddings=True, **c["tokens"],
    )


def build_model(name: str = "tiny", seed: int = 42, dtype=torch.float32, perturb: bool = True, round_bf16: bool = True):
    """Seeded random-init HF model (eager attention).  HF default init = N(0, 0.02) weights, unit norms,
    zero biases; with ``perturb`` the biases and norm weights are additionally randomised (seeded) so that
    the bias / norm-weight code paths are actually exercised by parity tests."""
    from transformers import Qwen2_5_VLForConditionalGeneration

    cfg = make_config(name)
    torch.manual_seed(seed)
    model = Qwen2_5_VLForConditionalGeneration._from_config(cfg, attn_implementation="eager", dtype=torch.float32)
    if perturb:
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for pname, p in model.named_parameters():
                if pname.endswith(".bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                elif p.dim() == 1:  # RMSNorm weights
                    p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
    if round_bf16:
        # A real checkpoint (Qwen2.5-VL / SocioReasoner-3B) stores bf16 values; round the random init to bf16-representable
        # numbers so the fp32 oracle and the bf16-storage CUDA path see the SAME weights (differences are then arithmetic only).
        with torch.no_grad():
            for p in model.parameters():
                p.copy_(p.to(torch.bfloat16).to(torch.float32))
    model = model.to(dtype).eval()
    return model


def build_model_fast(name: str = "3b", seed: int = 42, dtype=torch.float32):
    """Same architecture, weights tiled from one seeded N(0,0.02) chunk (norm weights 1): constructs the 3B model in
    seconds instead of minutes.  ONLY for timing the CPU baseline (values do not affect CPU time); parity tests use
    ``build_model``."""
    from transformers import Qwen2_5_VLForConditionalGeneration

    cfg = make_config(name)
    with torch.device("meta"):
        model = Qwen2_5_VLForConditionalGeneration._from_config(cfg, attn_implementation="eager", dtype=dtype)
    model = model.to_empty(device="cpu")
    g = torch.Generator().manual_seed(seed)
    chunk = (torch.randn(1 << 22, generator=g) * 0.02).to(dtype)
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if p.dim() == 1 and not pname.endswith(".bias"):
                p.fill_(1.0)
            else:
                flat = p.view(-1)
                n = flat.numel()
                reps = (n + chunk.numel() - 1) // chunk.numel()
                flat.copy_(chunk.repeat(reps)[:n] if reps > 1 else chunk[:n])
        vc, tc = cfg.vision_config, cfg.text_config
        hd = vc.hidden_size // vc.num_heads // 2
        model.model.visual.rotary_pos_emb.inv_freq.copy_(1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float) / hd)))
        thd = tc.hidden_size // tc.num_attention_heads
        theta = tc.rope_parameters["rope_theta"]
        inv = 1.0 / (theta ** (torch.arange(0, thd, 2, dtype=torch.float) / thd))
        model.model.language_model.rotary_emb.inv_freq.copy_(inv)
        if hasattr(model.model.language_model.rotary_emb, "original_inv_freq"):
            model.model.language_model.rotary_emb.original_inv_freq.copy_(inv)
    model.tie_weights()
    return model.eval()


def state_dict_4x_names(model) -> Dict[str, torch.Tensor]:
    """State dict renamed to the transformers-4.x checkpoint layout, which is the
    reference's checkpoint contract: ``visual.*``, ``model.layers.*``, ``model.embed_tokens``, ``lm_head``."""
    out = {}
    for k, v in model.state_dict().items():
        if k.startswith("model.visual."):
            out["visual." + k[len("model.visual."):]] = v
        elif k.startswith("model.language_model."):
            out["model." + k[len("model.language_model."):]] = v
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------------------------
# Synthetic inputs.  The image processor is HF's own (what the reference's collator
# calls: roll/datasets/collator.py:444-565 -> AutoProcessor), so preprocessing here IS the reference's.
# ----------------------------------------------------------------------------------------------
def synth_tile_uint8(i: int, side: int) -> np.ndarray:
    g = torch.Generator().manual_seed(1000 + i)
    return torch.randint(0, 256, (3, side, side), dtype=torch.uint8, generator=g).numpy()


def preprocess_tiles(tiles_chw_uint8: Sequence[np.ndarray], max_pixels: int = 768 * 768, min_pixels: int = 56 * 56):
    """HF Qwen2VLImageProcessor (bicubic smart_resize, rescale, normalise, T=2 duplicate, merge-block patchify)."""
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor

    ip = Qwen2VLImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels)
    pvs, thws = [], []
    for t in tiles_chw_uint8:
        img = Image.fromarray(np.ascontiguousarray(t.transpose(1, 2, 0)))
        out = ip(images=[img], return_tensors="pt")
        pvs.append(out["pixel_values"])
        thws.append(out["image_grid_thw"])
    return pvs, thws


def synth_prompt(i: int, merged_tokens, cfg_name: str = "3b", n_text: int = 128) -> torch.Tensor:
    """[im_start, user, \\n] + per image ([vision_start] + [image_pad]*M + [vision_end]) + n_text random ids + chat tail.
    ``merged_tokens``: int (one image) or a list (the reference's prompts carry two images, map + sat:
    roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py:74-81)."""
    tok = CONFIGS[cfg_name]["tokens"]
    vocab = CONFIGS[cfg_name]["text"]["vocab_size"]
    g = torch.Generator().manual_seed(2000 + i)
    if cfg_name == "3b":
        head, tail, hi = [151644, 872, 198], [151645, 198, 151644, 77091, 198], 151643
    else:
        head, tail, hi = [7001, 872, 198], [7002, 198, 7001, 6091, 198], 7000
    text = torch.randint(0, hi, (n_text,), generator=g).tolist()
    ms = [merged_tokens] if isinstance(merged_tokens, int) else list(merged_tokens)
    vis = []
    for m in ms:
        vis += [tok["vision_start_token_id"]] + [tok["image_token_id"]] * m + [tok["vision_end_token_id"]]
    ids = head + vis + text + tail
    assert max(ids) < vocab
    return torch.tensor(ids, dtype=torch.long)


@dataclass
class Batch:
    input_ids: torch.Tensor  # [B,S] (left-padded when ragged)
    attention_mask: torch.Tensor  # [B,S]
    position_ids: torch.Tensor  # [3,B,S]
    pixel_values: torch.Tensor  # [sum N, 1176]
    image_grid_thw: torch.Tensor  # [n_img, 3]


def make_batch(tile_ids: Sequence[int], sides: Sequence[int], cfg_name: str, max_pixels: int = 768 * 768,
               n_text=128, pad_to: Optional[int] = None, pad_token_id: int = 0, images_per_prompt: int = 1) -> Batch:
    """``images_per_prompt`` = k: prompt j carries tiles [j*k, (j+1)*k) (its text ids are seeded by its first tile id).
    ``n_text``: int, or one value per prompt (ragged batches)."""
    tiles = [synth_tile_uint8(i, s) for i, s in zip(tile_ids, sides)]
    pvs, thws = preprocess_tiles(tiles, max_pixels=max_pixels)
    tok = CONFIGS[cfg_name]["tokens"]
    prompts = []
    k = images_per_prompt
    assert len(tile_ids) % k == 0
    for j in range(len(tile_ids) // k):
        ms = [int(thw[0, 0] * thw[0, 1] * thw[0, 2]) // 4 for thw in thws[j * k:(j + 1) * k]]
        nt = n_text if isinstance(n_text, int) else n_text[j]
        prompts.append(synth_prompt(tile_ids[j * k], ms if k > 1 else ms[0], cfg_name, nt))
    S = max(len(p) for p in prompts) if pad_to is None else pad_to
    B = len(prompts)
    ids = torch.full((B, S), pad_token_id, dtype=torch.long)
    mask = torch.zeros((B, S), dtype=torch.long)
    for b, p in enumerate(prompts):  # left padding, as the reference's collator does (collator.py:421-565)
        ids[b, S - len(p):] = p
        mask[b, S - len(p):] = 1
    grid = torch.cat(thws, 0)
    pos, _ = host_ref.get_rope_index(ids, grid, mask, image_token_id=tok["image_token_id"],
                                     vision_start_token_id=tok["vision_start_token_id"])
    return Batch(ids, mask, pos, torch.cat(pvs, 0), grid)


# ----------------------------------------------------------------------------------------------
# Forward passes
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def vit_forward(model, pixel_values, image_grid_thw, capture_blocks: bool = False):
    """ViT -> (merged [sum M, 2048] in original order, pre-merger hidden [sum N,1280] in window order,
    optional per-block hidden states)."""
    vis = model.model.visual
    caps: List[torch.Tensor] = []
    hooks = []
    if capture_blocks:
        for blk in vis.blocks:
            hooks.append(blk.register_forward_hook(lambda m, i, o: caps.append(o.detach().clone())))
    out = vis(pixel_values.to(vis.dtype), grid_thw=image_grid_thw)
    for h in hooks:
        h.remove()
    return out.pooler_output, out.last_hidden_state, caps


@torch.no_grad()
def forward_logits(model, batch: Batch, output_hidden_states: bool = False):
    """The call of hf_strategy.py:84-90."""
    out = model(
        input_ids=batch.input_ids, attention_mask=batch.attention_mask, position_ids=batch.position_ids,
        use_cache=False, pixel_values=batch.pixel_values.to(model.dtype), image_grid_thw=batch.image_grid_thw,
        output_hidden_states=output_hidden_states,
    )
    return out


@torch.no_grad()
def greedy_decode(model, batch: Batch, steps: int, timers: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Greedy, EOS ignored, fixed ``steps`` new tokens.  Response positions follow the
    last prompt position + 1.. on all three axes (roll/utils/functionals.py:838-852).  Implemented as
    explicit KV-cache steps so the positions are the reference's, not HF-5.x generate()'s.
    Returns (new_tokens [B,steps], per-step fp32 logits of the chosen row [B,steps,V])."""
    import time as _time

    B, S = batch.input_ids.shape
    _t = _time.time()
    out = model(
        input_ids=batch.input_ids, attention_mask=batch.attention_mask, position_ids=batch.position_ids,
        use_cache=True, pixel_values=batch.pixel_values.to(model.dtype), image_grid_thw=batch.image_grid_thw,
    )
    if timers is not None:
        timers["prefill_s"] = _time.time() - _t
        timers["decode_s"] = []
    past = out.past_key_values
    logits = out.logits[:, -1].float()
    last_pos = batch.position_ids.max(dim=0)[0].max(dim=-1)[0]  # [B]
    mask = batch.attention_mask.clone()
    toks, all_logits = [], []
    for t in range(steps):
        nxt = logits.argmax(-1)
        toks.append(nxt)
        all_logits.append(logits)
        if t == steps - 1:
            break
        mask = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype)], 1)
        pos = (last_pos + 1 + t).view(1, B, 1).expand(3, B, 1)
        _t = _time.time()
        out = model(input_ids=nxt.view(B, 1), attention_mask=mask, position_ids=pos, use_cache=True,
                    past_key_values=past)
        if timers is not None:
            timers["decode_s"].append(_time.time() - _t)
        past = out.past_key_values
        logits = out.logits[:, -1].float()
    return torch.stack(toks, 1), torch.stack(all_logits, 1)


# Synthetic SAM-style epilogue inputs
def synth_mask_inputs(i: int, n_obj: int = 3, low: int = 256, gt_side: int = 768):
    g = torch.Generator().manual_seed(3000 + i)
    logits = torch.randn(n_obj, 3, low, low, generator=g)
    scores = torch.rand(n_obj, 3, generator=g)
    g2 = torch.Generator().manual_seed(4000 + i)
    gt = (torch.rand(gt_side, gt_side, generator=g2) < 0.3).to(torch.uint8).numpy() * 255
    return logits, scores, gt
