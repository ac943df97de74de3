"""The subset of ``roll/distributed/scheduler/protocol.py`` (DataProto, :145-617) and
``roll/datasets/collator.py:13-40`` (``collate_fn_to_dict_list``) that the strategies on the hot path use.
`tensordict` is not a dependency here: ``batch`` is a plain dict of tensors that share dim 0, with ``batch_size``."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch


class _Batch(dict):
    """dict of tensors with a TensorDict-like ``batch_size`` (only dim 0 is tracked)."""

    @property
    def batch_size(self):
        for v in self.values():
            return torch.Size([v.shape[0]])
        return torch.Size([0])

    def to(self, device):
        return _Batch({k: v.to(device) for k, v in self.items()})

    def clone(self):
        return _Batch({k: v.clone() for k, v in self.items()})


def collate_fn_to_dict_list(data_list: List[dict]) -> dict:
    """list[dict] -> dict: tensors concatenated on dim 0, everything else into object arrays (collator.py:13-40)."""
    tensors: Dict[str, list] = {}
    non_tensors: Dict[str, list] = {}
    for data in data_list:
        for key, val in data.items():
            (tensors if isinstance(val, torch.Tensor) else non_tensors).setdefault(key, []).append(val)
    out = {k: torch.cat(v, dim=0) for k, v in tensors.items()}
    for k, v in non_tensors.items():
        arr = np.empty([len(v)], dtype=object)
        arr[:] = v
        out[k] = arr
    return out


@dataclass
class DataProto:
    batch: Optional[_Batch] = None
    non_tensor_batch: Dict[str, np.ndarray] = field(default_factory=dict)
    meta_info: Dict = field(default_factory=dict)

    def __post_init__(self):
        if self.batch is not None and not isinstance(self.batch, _Batch):
            self.batch = _Batch(self.batch)
        self.check_consistency()

    def check_consistency(self):  # protocol.py:223-241
        n = len(self)
        if self.batch is not None:
            for k, v in self.batch.items():
                assert v.shape[0] == n, f"tensor {k} has batch {v.shape[0]} != {n}"
        for k, v in self.non_tensor_batch.items():
            assert isinstance(v, np.ndarray) and v.dtype == object and len(v) == n, f"non-tensor {k} must be an object array of len {n}"

    def __len__(self):
        if self.batch is not None and len(self.batch):
            return self.batch.batch_size[0]
        if self.non_tensor_batch:
            return len(next(iter(self.non_tensor_batch.values())))
        return 0

    @classmethod
    def from_dict(cls, tensors: Dict[str, torch.Tensor], non_tensors: Optional[dict] = None, meta_info: Optional[dict] = None):
        nt = {}
        for k, v in (non_tensors or {}).items():
            arr = np.empty([len(v)], dtype=object)
            arr[:] = list(v)
            nt[k] = arr
        return cls(batch=_Batch(tensors), non_tensor_batch=nt, meta_info=meta_info or {})

    def to(self, device):
        """Moves tensors only; non_tensor_batch never moves (hf_strategy.py:78 comment)."""
        if self.batch is not None:
            self.batch = self.batch.to(device)
        return self

    def chunk(self, chunks: int) -> List["DataProto"]:
        """Split along dim 0 with ``np.array_split`` sizes (protocol.py:550-595); every chunk gets meta_info."""
        n = len(self)
        assert n >= chunks, f"batch_size {n} < chunks {chunks}"
        sizes = [len(b) for b in np.array_split(np.arange(n), chunks)]
        bounds = np.concatenate([[0], np.cumsum(sizes)])
        out = []
        for i in range(chunks):
            a, b = int(bounds[i]), int(bounds[i + 1])
            batch = _Batch({k: v[a:b].clone() for k, v in self.batch.items()}) if self.batch is not None else None
            nt = {k: v[a:b] for k, v in self.non_tensor_batch.items()}
            out.append(DataProto(batch=batch, non_tensor_batch=nt, meta_info=self.meta_info))
        return out

    @staticmethod
    def concat(data: List["DataProto"]) -> "DataProto":
        """protocol.py:597-617: tensors cat on dim 0, object arrays concatenated, meta_info of the first."""
        batches = [d.batch for d in data if d.batch is not None]
        new_batch = _Batch({k: torch.cat([b[k] for b in batches], dim=0) for k in batches[0]}) if batches else None
        keys = data[0].non_tensor_batch.keys()
        nt = {k: np.concatenate([d.non_tensor_batch[k] for d in data], axis=0) for k in keys}
        return DataProto(batch=new_batch, non_tensor_batch=nt, meta_info=data[0].meta_info)
