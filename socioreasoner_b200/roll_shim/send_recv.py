"""Receive side of the reference's bucketed weight transfer (``roll/utils/send_recv_utils.py:64-163``).

Wire format of one bucket: an int8 buffer of ``bucket_size`` bytes plus ``meta_infos = {name: {bucket_start, tensor_start,
save_bytes, tensor_meta}}``; a tensor larger than the remaining space continues in the next bucket (``tensor_start`` > 0).
``tensor_meta`` is a meta-device tensor, or, after the sender's ``SendBucketManager.meta_to_dict``, a dict with
``shape`` / ``dtype``.  Only the receiver lives here (the sender belongs to the train strategy, out of scope)."""
from __future__ import annotations

from typing import Dict

import torch


def _meta_shape_dtype(meta):
    if isinstance(meta, dict):
        return tuple(meta["shape"]), meta["dtype"]
    return tuple(meta.shape), meta.dtype


def tensor_nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


class RecvBucketManager:
    """``send_recv_utils.py:137-163``: accumulates bucket slices into whole tensors; ``process_bucket`` returns the tensors
    that became complete with this bucket."""

    def __init__(self):
        self.waiting_tensors: Dict[str, torch.Tensor] = {}

    def process_bucket(self, tensors_meta: dict, buffer: torch.Tensor) -> Dict[str, torch.Tensor]:
        buf = buffer.view(-1).view(torch.int8)
        finished = {}
        for name, meta in tensors_meta.items():
            bucket_start, tensor_start, save_bytes = int(meta["bucket_start"]), int(meta["tensor_start"]), int(meta["save_bytes"])
            t = self.waiting_tensors.get(name)
            if t is None:
                shape, dtype = _meta_shape_dtype(meta["tensor_meta"])
                t = torch.empty(shape, dtype=dtype, device=buffer.device)
                self.waiting_tensors[name] = t
            t.view(-1).view(torch.int8)[tensor_start: tensor_start + save_bytes].copy_(buf[bucket_start: bucket_start + save_bytes])
            if tensor_start + save_bytes == tensor_nbytes(t):
                finished[name] = self.waiting_tensors.pop(name)
        return finished

    def clear(self):
        assert len(self.waiting_tensors) == 0, f"incomplete tensors left: {sorted(self.waiting_tensors)}"
