"""B200-native SAM2 image path: Hiera encoder + FPN neck -> prompt encoder -> two-way mask decoder, feeding
``sr_mask_epilogue`` -- what ``SegInferStrategy.segment`` drives through ``SAM2ImagePredictor.set_image / predict``
(roll/distributed/strategy/seg_strategy.py:26-72, roll/models/model_providers.py:515-562).  The arithmetic lives in the external
``sam2`` package (not in the reference tree, version unpinned); parity is pinned on its HF port
(``transformers.models.sam2``, the structural oracle SURVEY.md 8c names)."""
from .config import Sam2Cfg
from .engine import Sam2Engine
from .predictor import B200Sam2ImagePredictor

__all__ = ["Sam2Cfg", "Sam2Engine", "B200Sam2ImagePredictor"]
