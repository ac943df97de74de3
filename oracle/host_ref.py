"""Oracle restatements of the small in-tree algorithms on the hot path (TEST INFRASTRUCTURE).

Every function cites the reference file:line it follows (paths relative to /root/reference).
Pure numpy / torch on CPU; no CUDA, no imports from the product package.

Pinned: tests/test_reference_fixtures.py checks these restatements bit-exactly against tests/golden/reference_intree.pt --
outputs of the reference's OWN function sources, extracted with `ast` and executed by tools/make_reference_fixtures.py
(get_rope_index, log-prob / entropy + masking wrappers, postprocess_generate, compute_giou, the segment() epilogue around a
stub predictor, collate, the bucket wire format).  Not pinnable here: SAM2's bilinear-upsample + threshold tail (external
`sam2`, no source in the tree) -- mask_epilogue() restates it with the same torch call on synthetic low-res logits.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# M-RoPE position ids -- reference semantics (NOT transformers 5.x semantics).
# Follows mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py:319-441 and its
# call site roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py:330-368.
# Image temporal index is 0 (second_per_grid_t = 0 for images, :368-376), i.e. t = text_len + st_idx.
# ----------------------------------------------------------------------------------------------
def get_rope_index(
    input_ids: torch.Tensor,
    image_grid_thw: Optional[torch.Tensor],
    attention_mask: Optional[torch.Tensor] = None,
    *,
    spatial_merge_size: int = 2,
    image_token_id: int = 151655,
    vision_start_token_id: int = 151652,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (position_ids [3,B,S] int64, rope_deltas [B,1])."""
    if image_grid_thw is None:
        # text-only branch, reference :421-441
        if attention_mask is not None:
            position_ids = attention_mask.long().cumsum(-1) - 1
            position_ids.masked_fill_(attention_mask == 0, 1)
            position_ids = position_ids.unsqueeze(0).expand(3, -1, -1)
            mx = position_ids.max(0)[0].max(-1, keepdim=True)[0]
            return position_ids.clone(), mx + 1 - attention_mask.shape[-1]
        B, S = input_ids.shape
        return (
            torch.arange(S).view(1, 1, -1).expand(3, B, -1).clone(),
            torch.zeros(B, 1, dtype=input_ids.dtype),
        )
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    B, S = input_ids.shape
    position_ids = torch.ones(3, B, S, dtype=input_ids.dtype)  # padded slots keep 1, reference :338-345
    deltas = []
    image_index = 0
    for i in range(B):
        ids = input_ids[i][attention_mask[i] == 1]
        toks = ids.tolist()
        starts = torch.argwhere(ids == vision_start_token_id).squeeze(1)
        image_nums = int((ids[starts + 1] == image_token_id).sum()) if len(starts) else 0
        pos_list: List[torch.Tensor] = []
        st = 0
        for _ in range(image_nums):
            ed = toks.index(image_token_id, st)
            t, h, w = (int(x) for x in image_grid_thw[image_index])
            image_index += 1
            gh, gw = h // spatial_merge_size, w // spatial_merge_size
            text_len = ed - st
            st_idx = int(pos_list[-1].max()) + 1 if pos_list else 0
            pos_list.append(torch.arange(text_len).view(1, -1).expand(3, -1) + st_idx)
            t_index = torch.zeros(t * gh * gw, dtype=torch.long)  # images: second_per_grid_t == 0
            h_index = torch.arange(gh).view(1, -1, 1).expand(t, -1, gw).flatten()
            w_index = torch.arange(gw).view(1, 1, -1).expand(t, gh, -1).flatten()
            pos_list.append(torch.stack([t_index, h_index, w_index]) + text_len + st_idx)
            st = ed + t * gh * gw
        if st < len(toks):
            st_idx = int(pos_list[-1].max()) + 1 if pos_list else 0
            pos_list.append(torch.arange(len(toks) - st).view(1, -1).expand(3, -1) + st_idx)
        llm_positions = torch.cat(pos_list, dim=1).reshape(3, -1)
        position_ids[:, i, attention_mask[i] == 1] = llm_positions
        deltas.append(int(llm_positions.max()) + 1 - S)
    return position_ids, torch.tensor(deltas).unsqueeze(1)


# ----------------------------------------------------------------------------------------------
# M-RoPE table and application -- mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py:14-75
# (= HF:595-608, 627-669).  freqs[axis] = position_ids[axis] * inv_freq; the 64 frequency slots are taken from the
# temporal / height / width axis by mrope_section (16 | 24 | 24); x' = x * cos + rotate_half(x) * sin with the
# 64-slot table repeated over both halves of the 128-wide head.
# ----------------------------------------------------------------------------------------------
def mrope_cos_sin(position_ids: torch.Tensor, inv_freq: torch.Tensor, mrope_section: Sequence[int] = (16, 24, 24)):
    """position_ids [3, B, S] -> (cos, sin) fp32 [B, S, 64]: the section-selected half table sr_mrope_table produces."""
    fr = position_ids.float()[..., None] * inv_freq.float()[None, None, None, :]          # [3, B, S, 64]
    parts, lo = [], 0
    for axis, n in enumerate(mrope_section):
        parts.append(fr[axis % 3, :, :, lo: lo + n])
        lo += n
    sel = torch.cat(parts, -1)
    return sel.cos(), sel.sin()


def apply_mrope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B, S, H, 128], cos / sin [B, S, 64] -> rotated x (rotate-half convention)."""
    c = torch.cat([cos, cos], -1)[:, :, None, :]
    s = torch.cat([sin, sin], -1)[:, :, None, :]
    half = x.shape[-1] // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], -1)
    return x * c + rot * s


# ----------------------------------------------------------------------------------------------
# log-prob / entropy from logits -- roll/utils/functionals.py:193-205 and the masking wrappers
# roll/distributed/strategy/strategy.py:123-138.
# ----------------------------------------------------------------------------------------------
def log_probs_from_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    lp = F.log_softmax(logits.float(), dim=-1)
    return lp.gather(dim=-1, index=labels.unsqueeze(-1)).squeeze(-1)


def entropy_from_logits(logits: torch.Tensor) -> torch.Tensor:
    logits = logits.float()
    pd = F.softmax(logits, dim=-1)
    return torch.logsumexp(logits, dim=-1) - torch.sum(pd * logits, dim=-1)


def op_compute_log_probs(logits, input_ids, attention_mask):
    labels = input_ids[:, 1:].clone()
    labels[attention_mask[:, 1:] == 0] = 0
    return log_probs_from_logits(logits[:, :-1], labels) * attention_mask[:, 1:]


def op_compute_entropy(logits, attention_mask):
    return entropy_from_logits(logits)[:, :-1] * attention_mask[:, 1:]


# ----------------------------------------------------------------------------------------------
# Mask epilogue -- roll/distributed/strategy/seg_strategy.py:37-69.  SAM2's predictor (external
# `sam2`, un-vendored, not installed) ends with bilinear up-sampling of the low-res mask logits to the
# image size and a `> mask_threshold (0.0)` test; that tail is restated here with the same torch call
# (F.interpolate bilinear, align_corners=False) on SYNTHETIC low-res logits.
# ----------------------------------------------------------------------------------------------
def mask_epilogue(
    low_res_logits: torch.Tensor,  # [n_obj, 3, h, w] float32
    scores: torch.Tensor,  # [n_obj, 3] float32
    image_size: int = 756,
    out_size: int = 768,
) -> np.ndarray:
    import cv2

    if low_res_logits.shape[0] == 0:  # seg_strategy.py:40-45 -- empty prompt list
        return np.zeros((out_size, out_size), dtype=np.uint8)
    mask = np.zeros((image_size, image_size), dtype=np.uint8)
    for o in range(low_res_logits.shape[0]):
        up = F.interpolate(
            low_res_logits[o : o + 1].float(), (image_size, image_size), mode="bilinear", align_corners=False
        )[0]
        pred_masks = (up > 0.0).numpy()
        best = pred_masks[int(np.argmax(scores[o].numpy()))]  # seg_strategy.py:59
        mask = np.logical_or(mask, best).astype(np.uint8)  # seg_strategy.py:60
    return cv2.resize(mask, (out_size, out_size), interpolation=cv2.INTER_NEAREST)  # :65


def compute_giou(pred_mask: np.ndarray, gt_mask: np.ndarray) -> float:
    """roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py:45-58 (both-empty -> 1.0)."""
    p = pred_mask > 0
    g = gt_mask > 0
    inter = np.logical_and(p, g).sum()
    union = np.logical_or(p, g).sum()
    if union == 0:
        return 1.0
    return float(inter / union)


def compute_iou_reward(pred_mask: np.ndarray, gt_mask: np.ndarray) -> float:
    """roll/pipeline/rlvr/rewards/socioseg_rule_reward_worker.py:181-199 (empty union / shape mismatch -> 0.0)."""
    if pred_mask.shape != gt_mask.shape:
        return 0.0
    p = pred_mask > 0
    g = gt_mask > 0
    union = np.logical_or(p, g).sum()
    if union == 0:
        return 0.0
    return float(np.logical_and(p, g).sum() / union)


# ----------------------------------------------------------------------------------------------
# DP sharding of whole samples -- roll/distributed/scheduler/protocol.py:550-617 (chunk sizes follow
# np.array_split) and decorator.py:106-180 (dispatch by dp rank, concat in rank order on collect).
# ----------------------------------------------------------------------------------------------
def dp_chunk_indices(batch: int, dp_size: int) -> List[np.ndarray]:
    return [np.asarray(c) for c in np.array_split(np.arange(batch), dp_size)]


# ----------------------------------------------------------------------------------------------
# postprocess_generate -- roll/utils/functionals.py:768-872 (the part the strategy contract needs:
# pad to sequence_length, split prompt/response, masks).  M-RoPE response positions follow the last
# prompt position + 1.. (:838-852).
# ----------------------------------------------------------------------------------------------
def pad_generate_output(output: torch.Tensor, sequence_length: int, pad_token_id: int) -> torch.Tensor:
    if output.shape[1] >= sequence_length:
        return output[:, :sequence_length]
    return F.pad(output, (0, sequence_length - output.shape[1]), value=pad_token_id)


def postprocess_generate_ref(input_ids: torch.Tensor, attention_mask: torch.Tensor, position_ids: torch.Tensor,
                             output: torch.Tensor, num_return_sequences: int, sequence_length: int, pad_token_id: int):
    """What ``ActorWorker.generate`` hands to ``compute_log_probs`` (roll/utils/functionals.py:768-872, M-RoPE branch):
    the generated ``[b*n, S_prompt + S_new]`` ids are padded to ``sequence_length``; every row is shifted left by its prompt's
    left padding (-> right padding), ``response_mask`` marks the non-pad response tokens, ``attention_mask`` the valid prefix,
    and the 3-axis position ids continue with last prompt position + 1, + 2, ... on all axes and are shifted the same way.
    Returns a dict with the reference's keys (restated row by row: this is the oracle for the layout the fused log-prob path
    has to accept, not product code)."""
    out = pad_generate_output(output.clone(), sequence_length, pad_token_id)
    n, S_p = num_return_sequences, input_ids.shape[1]
    R = out.shape[0]
    assert R == input_ids.shape[0] * n and position_ids.dim() == 3
    prompts, responses = out[:, :S_p].clone(), out[:, S_p:].clone()
    am = attention_mask.repeat_interleave(n, dim=0)
    resp_mask = (responses != pad_token_id).to(am.dtype)
    pos = position_ids.repeat_interleave(n, dim=0)                                            # [R, 3, S_p]
    delta = torch.arange(1, sequence_length - S_p + 1).view(1, 1, -1)
    pos = torch.cat([pos, pos[..., -1:] + delta], dim=-1)                                     # [R, 3, sequence_length]
    new_am = torch.zeros(R, sequence_length, dtype=am.dtype)
    new_rm = torch.zeros_like(new_am)
    for i in range(R):
        shift = int(am[i].float().argmax())                       # left padding of the prompt
        valid = int(am[i].sum() + resp_mask[i].sum())
        n_resp = int(resp_mask[i].sum())
        if shift > 0:
            out[i, :-shift] = out[i, shift:].clone()
            pos[i, :, :-shift] = pos[i, :, shift:].clone()
            if S_p > n_resp:
                out[i, -shift:] = pad_token_id
        new_am[i, :valid] = 1
        new_rm[i, valid - n_resp: valid] = 1
    return {"prompts": prompts, "responses": responses, "input_ids": out, "attention_mask": new_am, "position_ids": pos,
            "prompt_mask": (new_am == 1) & (new_rm == 0), "response_mask": new_rm}


# ----------------------------------------------------------------------------------------------
# Sampling decode -- the reference's rollout parameters (examples/infer/rlvr_megatron.yaml:89-95: temperature 1.0,
# top_p 0.8, top_k 100 -> vLLM SamplingParams, roll/distributed/strategy/vllm_strategy.py:289-309).  Restated with the HF
# warper order temperature -> top-k -> top-p (TopPLogitsWarper removes tokens whose ASCENDING cumulative probability is
# <= 1 - top_p, i.e. keeps descending token j iff the mass strictly before it is < top_p) and an inverse-CDF pick from a
# given uniform so the result is deterministic.  float64.
# ----------------------------------------------------------------------------------------------
def sample_topk_topp_ref(logits: torch.Tensor, u: float, temperature: float, top_k: int, top_p: float):
    """logits [V] -> (token, (cdf_lo, cdf_hi) of that token in the renormalised kept distribution, kept token ids)."""
    x = logits.double().numpy()
    order = np.lexsort((np.arange(len(x)), -x))[: min(top_k, len(x))]     # value desc, index asc
    z = x[order] / temperature
    p = np.exp(z - z.max())
    p /= p.sum()
    excl = np.cumsum(p) - p
    keep = excl < top_p
    keep[0] = True
    q = p[keep] / p[keep].sum()
    cdf = np.cumsum(q)
    j = int(np.searchsorted(cdf, u, side="right"))
    j = min(j, len(q) - 1)
    return int(order[keep][j]), (float(cdf[j] - q[j]), float(cdf[j])), order[keep]


# ----------------------------------------------------------------------------------------------
# Sender side of the bucketed weight transfer -- roll/utils/send_recv_utils.py:64-135 (TensorBucket,
# SendBucketManager) as driven by roll/distributed/strategy/megatron_strategy.py:411-448.  Test infrastructure for
# the receive side in socioreasoner_b200/roll_shim/send_recv.py: produces the exact wire format
# (int8 bucket buffer + {name: {bucket_start, tensor_start, save_bytes, tensor_meta}}).
# ----------------------------------------------------------------------------------------------
class SendBucketManager:
    def __init__(self, bucket_size: int, device="cpu"):
        self.bucket_size = bucket_size
        self.buffer = torch.empty(bucket_size, dtype=torch.int8, device=device)
        self.write_index = 0
        self.tensors_meta = {}

    def _push_part(self, tensor, tensor_start, name):
        required = tensor.numel() * tensor.element_size() - tensor_start
        bucket_start = self.write_index
        save_bytes = min(required, self.bucket_size - bucket_start)
        self.buffer[bucket_start: bucket_start + save_bytes].copy_(tensor.view(-1).view(torch.int8)[tensor_start: tensor_start + save_bytes])
        self.tensors_meta[name] = {"bucket_start": bucket_start, "tensor_start": tensor_start, "save_bytes": save_bytes,
                                   "tensor_meta": torch.empty_like(tensor, device="meta")}
        self.write_index += save_bytes
        return save_bytes

    def push_tensor(self, tensor: torch.Tensor, name: str):
        """Generator: yields (tensors_meta, buffer) every time the bucket fills up (send_recv_utils.py:112-120)."""
        tensor_start, required = 0, tensor.numel() * tensor.element_size()
        while tensor_start < required:
            tensor_start += self._push_part(tensor, tensor_start, name)
            if self.write_index == self.bucket_size:
                yield self.tensors_meta, self.buffer
                self.tensors_meta, self.write_index = {}, 0

    def pop_last_bucket(self):
        if self.write_index > 0:
            return self.tensors_meta, self.buffer
        return None, None

    @staticmethod
    def meta_to_dict(meta_infos):
        """send_recv_utils.py:127-141: meta tensors -> plain dicts (what crosses the RPC boundary)."""
        for meta_info in meta_infos.values():
            t = meta_info["tensor_meta"]
            meta_info["tensor_meta"] = {"shape": list(t.shape), "dtype": t.dtype, "layout": t.layout, "device": t.device}
        return meta_infos


def send_buckets(named_tensors, bucket_size: int, as_dict: bool = True):
    """All buckets of one model update, in order: list of (meta_infos, buffer copy)."""
    mgr, out = SendBucketManager(bucket_size), []
    emit = lambda meta, buf: out.append((SendBucketManager.meta_to_dict(dict(meta)) if as_dict else dict(meta), buf.clone()))
    for name, t in named_tensors:
        for meta, buf in mgr.push_tensor(t.contiguous(), name):
            emit(meta, buf)
    meta, buf = mgr.pop_last_bucket()
    if meta is not None:
        emit(meta, buf)
    return out


# ----------------------------------------------------------------------------------------------
# Image path in front of the ViT, exactly as the reference runs it on the CPU:
#   process_image (roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py:126-144): PIL resize to smart_resize(h, w) with the
#   processor's resample (bicubic), THEN the HF processor (roll/datasets/collator.py:456-462), whose own resize is the identity
#   on the pre-resized image.  Returns (pixel_values fp32 [sum N, 1176], image_grid_thw [n, 3], the resized uint8 images).
# ----------------------------------------------------------------------------------------------
def reference_image_path(images_hwc_uint8, min_pixels: int = 56 * 56, max_pixels: int = 768 * 768):
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor, smart_resize

    ip = Qwen2VLImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels)
    pvs, thws, resized = [], [], []
    for a in images_hwc_uint8:
        img = Image.fromarray(np.ascontiguousarray(a))
        rh, rw = smart_resize(img.height, img.width, factor=28, min_pixels=min_pixels, max_pixels=max_pixels)
        img = img.resize((rw, rh), resample=Image.BICUBIC)
        resized.append(np.asarray(img))
        out = ip(images=[img], return_tensors="pt")
        pvs.append(out["pixel_values"])
        thws.append(out["image_grid_thw"])
    return torch.cat(pvs, 0), torch.cat(thws, 0), resized
