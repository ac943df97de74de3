"""SAM2 geometry (Hiera trunk + FPN neck + prompt encoder + mask decoder)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple


@dataclass
class Sam2Cfg:
    """Defaults = sam2-hiera-large (sam2_hiera_l.yaml of the ``sam2`` package: the checkpoint the reference loads,
    roll/models/model_providers.py:540-545)."""
    image_size: int = 1024
    embed_dims: Tuple[int, ...] = (144, 288, 576, 1152)
    num_heads: Tuple[int, ...] = (2, 4, 8, 16)
    blocks_per_stage: Tuple[int, ...] = (2, 6, 36, 4)
    global_attention_blocks: Tuple[int, ...] = (23, 33, 43)
    window_sizes: Tuple[int, ...] = (8, 4, 16, 8)
    pos_embed_bkg: Tuple[int, int] = (7, 7)
    patch_kernel: int = 7
    patch_stride: int = 4
    patch_padding: int = 3
    mlp_ratio: float = 4.0
    layer_norm_eps: float = 1e-6
    fpn_hidden: int = 256
    fpn_top_down_levels: Tuple[int, ...] = (2, 3)
    # decoder
    dec_hidden: int = 256
    dec_heads: int = 8
    dec_layers: int = 2
    dec_mlp: int = 2048
    dec_downsample: int = 2
    iou_hidden: int = 256
    iou_depth: int = 3
    num_multimask: int = 3
    prompt_patch: int = 16

    def validate(self) -> None:
        r = self.image_size // self.patch_stride
        assert self.image_size % (self.patch_stride * 8) == 0
        for s, ws in enumerate(self.window_sizes):
            assert (r >> s) % ws == 0, f"stage {s}: resolution {r >> s} is not a multiple of the window size {ws} (padding is not supported)"
            assert self.embed_dims[s] % self.num_heads[s] == 0 and self.embed_dims[s] // self.num_heads[s] <= 128
            if s > 0:
                assert self.window_sizes[s - 1] % 2 == 0
        assert self.image_size // self.prompt_patch == r >> 2, "image embedding = stage-3 resolution"
        assert self.embed_dims[0] % 8 == 0

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_stride

    @classmethod
    def from_hf(cls, hf_cfg) -> "Sam2Cfg":
        v, b = hf_cfg.vision_config, hf_cfg.vision_config.backbone_config
        d, p = hf_cfg.mask_decoder_config, hf_cfg.prompt_encoder_config
        isz = b.image_size[0] if isinstance(b.image_size, (list, tuple)) else b.image_size
        return cls(image_size=isz, embed_dims=tuple(b.embed_dim_per_stage), num_heads=tuple(b.num_attention_heads_per_stage),
                   blocks_per_stage=tuple(b.blocks_per_stage), global_attention_blocks=tuple(b.global_attention_blocks),
                   window_sizes=tuple(b.window_size_per_stage), pos_embed_bkg=tuple(b.window_positional_embedding_background_size),
                   patch_kernel=b.patch_kernel_size[0], patch_stride=b.patch_stride[0], patch_padding=b.patch_padding[0],
                   mlp_ratio=b.mlp_ratio, layer_norm_eps=b.layer_norm_eps, fpn_hidden=v.fpn_hidden_size,
                   fpn_top_down_levels=tuple(v.fpn_top_down_levels), dec_hidden=d.hidden_size, dec_heads=d.num_attention_heads,
                   dec_layers=d.num_hidden_layers, dec_mlp=d.mlp_dim, dec_downsample=d.attention_downsample_rate,
                   iou_hidden=d.iou_head_hidden_dim, iou_depth=d.iou_head_depth, num_multimask=d.num_multimask_outputs,
                   prompt_patch=p.patch_size)
