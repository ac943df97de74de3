"""Model configuration for the Qwen2.5-VL hot path.  Accepts the transformers-4.x flat layout the reference's
checkpoints ship (``rope_scaling{type: mrope, mrope_section}``, flat text fields; see
``mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py:98``) as well as the 5.x nested
``text_config`` / ``rope_parameters`` layout."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Tuple


@dataclass
class VisionConfig:
    depth: int = 32
    hidden_size: int = 1280
    intermediate_size: int = 3420
    num_heads: int = 16
    patch_size: int = 14
    spatial_merge_size: int = 2
    temporal_patch_size: int = 2
    window_size: int = 112
    out_hidden_size: int = 2048
    fullatt_block_indexes: Tuple[int, ...] = (7, 15, 23, 31)
    in_channels: int = 3
    rope_theta: float = 10000.0
    rms_norm_eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def patch_dim(self) -> int:
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size

    @property
    def intermediate_padded(self) -> int:
        return (self.intermediate_size + 127) // 128 * 128  # 3420 -> 3456: 16-byte rows for TMA + whole SwiGLU groups


@dataclass
class TextConfig:
    hidden_size: int = 2048
    intermediate_size: int = 11008
    num_hidden_layers: int = 36
    num_attention_heads: int = 16
    num_key_value_heads: int = 2
    vocab_size: int = 151936
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    mrope_section: Tuple[int, int, int] = (16, 24, 24)
    tie_word_embeddings: bool = True

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class SocioReasonerConfig:
    text: TextConfig = field(default_factory=TextConfig)
    vision: VisionConfig = field(default_factory=VisionConfig)
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_end_token_id: int = 151653

    @staticmethod
    def from_dict(d: Dict[str, Any]) -> "SocioReasonerConfig":
        t = dict(d.get("text_config") or {})
        flat = {k: v for k, v in d.items() if k not in ("text_config", "vision_config")}
        for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads",
                  "vocab_size", "rms_norm_eps", "rope_theta", "tie_word_embeddings", "rope_scaling", "rope_parameters"):
            if k not in t and k in flat:
                t[k] = flat[k]
        rp = t.get("rope_parameters") or t.get("rope_scaling") or {}
        text = TextConfig(
            hidden_size=t.get("hidden_size", 2048), intermediate_size=t.get("intermediate_size", 11008),
            num_hidden_layers=t.get("num_hidden_layers", 36), num_attention_heads=t.get("num_attention_heads", 16),
            num_key_value_heads=t.get("num_key_value_heads", 2), vocab_size=t.get("vocab_size", 151936),
            rms_norm_eps=t.get("rms_norm_eps", 1e-6),
            rope_theta=float(rp.get("rope_theta", t.get("rope_theta", 1000000.0))),
            mrope_section=tuple(rp.get("mrope_section", (16, 24, 24))),
            tie_word_embeddings=bool(d.get("tie_word_embeddings", t.get("tie_word_embeddings", True))),
        )
        v = dict(d.get("vision_config") or {})
        vision = VisionConfig(
            depth=v.get("depth", 32), hidden_size=v.get("hidden_size", 1280), intermediate_size=v.get("intermediate_size", 3420),
            num_heads=v.get("num_heads", 16), patch_size=v.get("patch_size", 14), spatial_merge_size=v.get("spatial_merge_size", 2),
            temporal_patch_size=v.get("temporal_patch_size", 2), window_size=v.get("window_size", 112),
            out_hidden_size=v.get("out_hidden_size", 2048),
            fullatt_block_indexes=tuple(v.get("fullatt_block_indexes", (7, 15, 23, 31))), in_channels=v.get("in_channels", v.get("in_chans", 3)),
        )
        return SocioReasonerConfig(
            text=text, vision=vision, image_token_id=d.get("image_token_id", 151655), video_token_id=d.get("video_token_id", 151656),
            vision_start_token_id=d.get("vision_start_token_id", 151652), vision_end_token_id=d.get("vision_end_token_id", 151653),
        )

    @staticmethod
    def from_hf(hf_config) -> "SocioReasonerConfig":
        return SocioReasonerConfig.from_dict(hf_config.to_dict())

    @staticmethod
    def from_pretrained(path: str) -> "SocioReasonerConfig":
        with open(os.path.join(path, "config.json")) as f:
            return SocioReasonerConfig.from_dict(json.load(f))

    def validate(self) -> None:
        v, t = self.vision, self.text
        if v.hidden_size != 1280 or v.num_heads != 16 or t.hidden_size != 2048 or t.head_dim != 128 or v.out_hidden_size != t.hidden_size:
            raise NotImplementedError(
                "kernels are specialised for the Qwen2.5-VL-3B geometry (ViT 1280/16 heads, decoder 2048/head_dim 128); "
                f"got vision {v.hidden_size}/{v.num_heads}, text {t.hidden_size}/{t.head_dim}")
        if sum(t.mrope_section) != t.head_dim // 2:
            raise ValueError("mrope_section must sum to head_dim/2")
