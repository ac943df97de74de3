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


# ----------------------------------------------------------------------------------------------------------------------
# GPU preprocessing: uint8 tile -> pixel_values rows, bit-exact with what the reference's workers receive.
# The reference resizes on the driver with PIL (``process_image``, rlvr_socioseg_vlm_pipeline_infer.py:126-144:
# ``image.resize((w, h), resample=BICUBIC)``) and then calls the HF processor, whose own resize is then the identity.  PIL's
# 8-bit resampler (libImaging/Resample.c, Pillow >= 7) is integer arithmetic on double-precision coefficient tables: the tables
# are restated here on the host (one small table per (in, out) size pair, cached); the two integer passes, the 1/255 rescale,
# the normalisation, the T = 2 frame duplication and the merge-block patch layout run in ``sr_preprocess_tile`` on the device.
# ----------------------------------------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2
IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)
_coeff_cache = {}


def _bicubic(x: np.ndarray) -> np.ndarray:
    """Pillow's bicubic kernel (a = -0.5), evaluated in float64 in the same operation order as Resample.c."""
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def pil_bicubic_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the bicubic filter (support 2) over the whole axis:
    returns (bounds int32 [out, 2] = (first input index, tap count), coeffs int32 [out, ksize] in 22-bit fixed point)."""
    key = (in_size, out_size)
    if key in _coeff_cache:
        return _coeff_cache[key]
    scale = np.float64(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = _bicubic((np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:             # sequential sum, as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fx = w * float(1 << PIL_PRECISION_BITS)
        kk[xx, :xmax] = np.where(w < 0, np.trunc(-0.5 + fx), np.trunc(0.5 + fx)).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    _coeff_cache[key] = (bounds, kk)
    return bounds, kk


def normalize_constants() -> Tuple[np.ndarray, np.ndarray]:
    """The fused rescale + normalise of the HF processor backend (``_fuse_mean_std_and_rescale_factor`` + torchvision
    ``normalize``): pixel = (float(u8) - mean_s) / std_s in fp32, with mean_s = fp32(mean) * fp32(1 / rescale_factor) and std_s
    likewise (fp32 products, rescale_factor = 1/255).  Verified bit-exact against the processor in tests/test_preprocess.py."""
    inv = np.float32(1.0 / (1.0 / 255))
    mean = np.array([np.float32(m) * inv for m in IMAGE_MEAN], dtype=np.float32)
    std = np.array([np.float32(v) * inv for v in IMAGE_STD], dtype=np.float32)
    return mean, std


class GpuPreprocessor:
    """uint8 tiles -> ``pixel_values`` / ``image_grid_thw`` on the device (``sr_preprocess_tile``), bit-exact with PIL resize +
    HF processor.  Coefficient tables are built once per (in, out) size pair and kept on the device."""

    def __init__(self, device="cuda", min_pixels: int = 56 * 56, max_pixels: int = REFERENCE_MAX_PIXELS, patch_size: int = 14,
                 merge_size: int = 2):
        self.device = torch.device(device)
        self.min_pixels, self.max_pixels, self.patch, self.merge = min_pixels, max_pixels, patch_size, merge_size
        self.mean, self.std = normalize_constants()
        self._tabs = {}
        self._scratch = None

    def _tab(self, n_in: int, n_out: int):
        if n_in == n_out:
            return None
        key = (n_in, n_out)
        if key not in self._tabs:
            b, k = pil_bicubic_coeffs(n_in, n_out)
            self._tabs[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(np.ascontiguousarray(k)).to(self.device))
        return self._tabs[key]

    def target_size(self, h: int, w: int) -> Tuple[int, int]:
        return smart_resize(h, w, factor=self.patch * self.merge, min_pixels=self.min_pixels, max_pixels=self.max_pixels)

    def __call__(self, tiles, out_dtype=torch.float32, out: torch.Tensor = None, row_maps=None):
        """tiles: sequence of uint8 [H, W, 3] tensors / arrays (host or device; host tensors are copied, pinned ones asynchronously).
        Returns (pixel_values [sum N, 1176], image_grid_thw int64 [n, 3] (host))."""
        from . import lib

        sizes, grids = [], []
        for t in tiles:
            h, w = int(t.shape[0]), int(t.shape[1])
            oh, ow = self.target_size(h, w)
            sizes.append((h, w, oh, ow))
            grids.append([1, oh // self.patch, ow // self.patch])
        n_rows = [g[1] * g[2] for g in grids]
        cols = 3 * 2 * self.patch * self.patch
        if out is None:
            out = torch.empty((sum(n_rows), cols), dtype=out_dtype, device=self.device)
        r0 = 0
        for i, (t, (h, w, oh, ow)) in enumerate(zip(tiles, sizes)):
            timg = t if torch.is_tensor(t) else torch.from_numpy(np.ascontiguousarray(t))
            timg = timg.to(self.device, non_blocking=True).contiguous()
            scratch = None
            if ow != w:
                need = h * ow * 3
                if self._scratch is None or self._scratch.numel() < need:
                    self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
                scratch = self._scratch
            lib.preprocess_tile(timg, oh, ow, self._tab(w, ow), self._tab(h, oh), scratch, self.mean, self.std, out[r0:r0 + n_rows[i]],
                                row_map=None if row_maps is None else row_maps[i], patch_size=self.patch)
            r0 += n_rows[i]
        return out, torch.tensor(grids, dtype=torch.int64)
