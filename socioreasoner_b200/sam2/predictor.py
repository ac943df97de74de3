"""``SAM2ImagePredictor``-compatible front end of the B200 SAM2 engine -- the object ``sam2_seg_model_provider`` returns in the
reference (roll/models/model_providers.py:515-562) and ``SegInferStrategy.segment`` drives (seg_strategy.py:47-58):

    predictor.set_image(PIL image)                                   # once per sample
    masks, scores, low_res_logits = predictor.predict(box=..., point_coords=..., point_labels=...)     # per object

Semantics follow the ``sam2`` package's predictor: the image is resized to the model's square input (bilinear) and normalised with
the ImageNet statistics; prompt coordinates are scaled from the original image to the model input; a box becomes two corner
points (labels 2, 3) in front of the click points, and a padding "not a point" token closes the prompt; the three multi-mask
candidates come back as low-res logits (clamped to +-32) plus, for API parity, thresholded masks at the original size."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .config import Sam2Cfg
from .engine import Sam2Engine

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class B200Sam2ImagePredictor:
    mask_threshold = 0.0

    def __init__(self, engine: Sam2Engine):
        self.engine, self.cfg = engine, engine.cfg
        self.model = engine            # the reference touches ``predictor.model`` for offload / eval switches only
        self._orig_hw: Optional[Tuple[int, int]] = None

    @classmethod
    def from_state_dict(cls, state_dict, cfg: Sam2Cfg, device="cuda"):
        from .ops_cuda import CudaOps
        from .weights import pack
        return cls(Sam2Engine(cfg, pack(state_dict, cfg, device), CudaOps(device)))

    @classmethod
    def from_hf_model(cls, hf_model, device="cuda"):
        return cls.from_state_dict(hf_model.state_dict(), Sam2Cfg.from_hf(hf_model.config), device)

    @classmethod
    def from_pretrained(cls, path: str, device="cuda"):
        """HF-format SAM2 checkpoint directory (config.json + safetensors with ``Sam2Model`` names)."""
        import json
        import os
        from safetensors import safe_open
        from transformers import Sam2Config
        cfg = Sam2Cfg.from_hf(Sam2Config.from_pretrained(path))
        sd = {}
        for f in sorted(x for x in os.listdir(path) if x.endswith(".safetensors")):
            with safe_open(os.path.join(path, f), framework="pt", device="cpu") as fh:
                for k in fh.keys():
                    sd[k] = fh.get_tensor(k)
        return cls.from_state_dict(sd, cfg, device)

    # ------------------------------------------------------------------ SAM2ImagePredictor API
    def set_image(self, image) -> None:
        arr = np.asarray(image.convert("RGB") if hasattr(image, "convert") else image, dtype=np.uint8)
        assert arr.ndim == 3 and arr.shape[2] == 3, "HWC RGB image expected"
        self._orig_hw = (arr.shape[0], arr.shape[1])
        px = self.engine.ops.preprocess_image(arr, self.cfg.image_size, IMAGENET_MEAN, IMAGENET_STD)     # fp32 [3, S, S] on the device
        self.engine.set_image(px)

    def _prompt(self, point_coords, point_labels, box):
        assert self._orig_hw is not None, "set_image first"
        h, w = self._orig_hw
        S = float(self.cfg.image_size)
        coords, labels = [], []
        if box is not None:
            b = np.asarray(box, dtype=np.float32).reshape(2, 2)
            coords += [[b[0, 0] / w * S, b[0, 1] / h * S], [b[1, 0] / w * S, b[1, 1] / h * S]]
            labels += [2, 3]
        if point_coords is not None:
            assert point_labels is not None, "point_labels must be supplied with point_coords"
            pc = np.asarray(point_coords, dtype=np.float32).reshape(-1, 2)
            for (x, y), l in zip(pc, np.asarray(point_labels).reshape(-1)):
                coords.append([x / w * S, y / h * S])
                labels.append(int(l))
        coords.append([0.0, 0.0])
        labels.append(-1)
        return np.asarray(coords, np.float32), np.asarray(labels, np.int64)

    def predict_logits(self, point_coords=None, point_labels=None, box=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device fast path: (low-res logits fp32 [3, 256, 256] clamped to +-32, scores fp32 [3]) -- feeds ``sr_mask_epilogue``."""
        coords, labels = self._prompt(point_coords, point_labels, box)
        logits, iou = self.engine.predict_logits(coords, labels)
        return logits.clamp_(-32.0, 32.0), iou

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output: bool = True,
                return_logits: bool = False, normalize_coords: bool = True):
        if mask_input is not None or not multimask_output:
            raise NotImplementedError("the SocioSeg path prompts with boxes / points and takes the 3 multi-mask candidates")
        logits, iou = self.predict_logits(point_coords, point_labels, box)
        h, w = self._orig_hw
        up = torch.nn.functional.interpolate(logits[None], size=(h, w), mode="bilinear", align_corners=False)[0]
        masks = up if return_logits else (up > self.mask_threshold)
        return masks.cpu().numpy(), iou.cpu().numpy(), logits.cpu().numpy()
