"""Host-side tile preprocessing (the collator's job in the reference, NOT on the GPU hot path).

The reference calls the HF processor per sample (``roll/datasets/collator.py:444-565`` ->
``AutoProcessor`` -> ``Qwen2VLImageProcessor``: bicubic ``smart_resize`` to multiples of 28 under ``max_pixels``,
rescale 1/255, normalise, duplicate the frame to T=2, patchify into 2x2-merge-block order ->
``pixel_values [N,1176]`` + ``image_grid_thw``).  This module drives the same processor so the tensors entering
the CUDA path are bit-identical to what the reference's workers receive.  ``max_pixels`` default 768*768 is the
reference-effective value (``rlvr_socioseg_vlm_pipeline_infer.py:518-521``: the yaml's 1344*1344 is silently dropped).
GPU-side preprocessing is a ranked "next" row.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

REFERENCE_MAX_PIXELS = 768 * 768
YAML_MAX_PIXELS = 1344 * 1344
_processors = {}


def _processor(max_pixels: int, min_pixels: int):
    key = (max_pixels, min_pixels)
    if key not in _processors:
        from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor

        _processors[key] = Qwen2VLImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels)
    return _processors[key]


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = REFERENCE_MAX_PIXELS):
    """HF ``image_processing_qwen2_vl.smart_resize`` (transformers/models/qwen2_vl/image_processing_qwen2_vl.py:62-88)."""
    import math

    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def preprocess_tiles(tiles_chw_uint8: Sequence[np.ndarray], max_pixels: int = REFERENCE_MAX_PIXELS,
                     min_pixels: int = 56 * 56) -> Tuple[torch.Tensor, torch.Tensor]:
    """uint8 [3,H,W] tiles -> (pixel_values fp32 [sum N, 1176], image_grid_thw int64 [n, 3])."""
    from PIL import Image

    ip = _processor(max_pixels, min_pixels)
    pvs, thws = [], []
    for t in tiles_chw_uint8:
        img = Image.fromarray(np.ascontiguousarray(np.asarray(t).transpose(1, 2, 0)))
        out = ip(images=[img], return_tensors="pt")
        pvs.append(out["pixel_values"])
        thws.append(out["image_grid_thw"])
    return torch.cat(pvs, 0), torch.cat(thws, 0)


# ---- synthetic workload of SURVEY.md 8d (seeds: tile 1000+i, prompt 2000+i, mask logits 3000+i, GT mask 4000+i) ----
def synth_tile(i: int, side: int = 1024) -> np.ndarray:
    g = torch.Generator().manual_seed(1000 + i)
    return torch.randint(0, 256, (3, side, side), dtype=torch.uint8, generator=g).numpy()


def synth_prompt(i: int, merged_tokens: int, image_token_id: int = 151655, vision_start: int = 151652, vision_end: int = 151653,
                 n_text: int = 128, text_vocab: int = 151643, head=(151644, 872, 198), tail=(151645, 198, 151644, 77091, 198)) -> torch.Tensor:
    g = torch.Generator().manual_seed(2000 + i)
    text = torch.randint(0, text_vocab, (n_text,), generator=g).tolist()
    return torch.tensor(list(head) + [vision_start] + [image_token_id] * merged_tokens + [vision_end] + text + list(tail), dtype=torch.long)


def synth_mask_inputs(i: int, n_obj: int = 3, low: int = 256, gt_side: int = 768):
    g = torch.Generator().manual_seed(3000 + i)
    logits = torch.randn(n_obj, 3, low, low, generator=g)
    scores = torch.rand(n_obj, 3, generator=g)
    g2 = torch.Generator().manual_seed(4000 + i)
    gt = (torch.rand(gt_side, gt_side, generator=g2) < 0.3).to(torch.uint8) * 255
    return logits, scores, gt


def left_pad(prompts: List[torch.Tensor], pad_token_id: int = 151643, pad_to: int = None):
    """Left padding as the reference's collator does (collator.py:421-565)."""
    S = max(len(p) for p in prompts) if pad_to is None else pad_to
    ids = torch.full((len(prompts), S), pad_token_id, dtype=torch.long)
    mask = torch.zeros((len(prompts), S), dtype=torch.long)
    for b, p in enumerate(prompts):
        ids[b, S - len(p):] = p
        mask[b, S - len(p):] = 1
    return ids, mask
