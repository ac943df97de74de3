"""HF-style module surface over the B200 engine -- what the reference's strategies and model providers touch.

The reference builds its actor through the HF Auto surface (``roll/models/model_providers.py:106-127``:
``AutoConfig.from_pretrained`` + ``AutoModelForVision2Seq.from_pretrained``) and then only ever calls

    model(input_ids=[b,S], attention_mask=[b,S], position_ids=[3,b,S], use_cache=False,
          pixel_values=[sum N,1176], image_grid_thw=[k,3], force_vit_image=True).logits     (hf_strategy.py:84-91)
    model.generate(input_ids=, attention_mask=, use_cache=True, pixel_values=, image_grid_thw=, **generation_config)
                                                                                             (hf_strategy.py:123-126)

This class answers exactly those calls (same argument names / meaning / error behaviour) on the hand-written
kernels.  It is inference-only (``requires_grad`` nowhere); ``state_dict`` key names follow the checkpoint contract
(SURVEY.md Appendix A).
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from . import lib
from .config import SocioReasonerConfig
from .engine import Engine
from .weights import PackedWeights, pack, random_state_dict


class CausalLMOutput(SimpleNamespace):
    """Minimal stand-in for HF's ``Qwen2_5_VLCausalLMOutputWithPast`` (only ``.logits`` is read by the reference,
    ``hf_strategy.py:91``; ``rope_deltas`` kept for API familiarity)."""


class SocioReasonerB200ForConditionalGeneration:
    """Drop-in for ``Qwen2_5_VLForConditionalGeneration`` on the reference's inference path."""

    def __init__(self, config: SocioReasonerConfig, weights: PackedWeights, device="cuda", logits_dtype=torch.float32):
        self.config = config
        self.device = torch.device(device)
        self.engine = Engine(config, weights, device)
        self.logits_dtype = logits_dtype
        self.training = False

    # ---- construction -------------------------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, state_dict: Dict[str, torch.Tensor], config, device="cuda", **kw):
        cfg = config if isinstance(config, SocioReasonerConfig) else SocioReasonerConfig.from_hf(config)
        return cls(cfg, pack(state_dict, cfg, device), device, **kw)

    @classmethod
    def from_hf_model(cls, hf_model, device="cuda", **kw):
        return cls.from_state_dict(hf_model.state_dict(), hf_model.config, device, **kw)

    @classmethod
    def random_init(cls, config: SocioReasonerConfig, seed: int = 42, device="cuda", **kw):
        sd = random_state_dict(config, seed, device)
        w = pack(sd, config, device)
        del sd
        torch.cuda.empty_cache()
        return cls(config, w, device, **kw)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, device="cuda", **kw):
        """Loads an HF checkpoint directory (config.json + *.safetensors), the surface
        ``model_providers.py:123-127`` uses.  Unknown HF kwargs (torch_dtype, attn_implementation, ...) are ignored:
        there is exactly one implementation here (bf16 storage, fp32 accumulate, tcgen05 kernels)."""
        path = pretrained_model_name_or_path
        cfg = SocioReasonerConfig.from_pretrained(path)
        from safetensors import safe_open  # transformers dependency, present wherever HF checkpoints are

        sd: Dict[str, torch.Tensor] = {}
        files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
        if not files:
            raise FileNotFoundError(f"no *.safetensors under {path}")
        for f in files:
            with safe_open(os.path.join(path, f), framework="pt", device="cpu") as fh:
                for k in fh.keys():
                    sd[k] = fh.get_tensor(k)
        return cls.from_state_dict(sd, cfg, device, **{k: v for k, v in kw.items() if k in ("logits_dtype",)})

    def load_weights(self, named_tensors) -> List[str]:
        """In-place update from ``(hf_name, tensor)`` pairs -- the call the reference's weight sync ends in
        (``roll/third_party/vllm/worker_helper.py:104-116``).  Pointers stay stable: captured decode graphs remain valid."""
        from .weights import update_weights
        return update_weights(self.engine.w, named_tensors, self.config)

    # ---- nn.Module-ish surface ------------------------------------------------------------------------
    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("inference-only hot path (training backward is out of scope, SURVEY.md 8f)")
        return self

    def to(self, *a, **k):
        return self

    @property
    def dtype(self):
        return torch.bfloat16

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    # ---- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, use_cache: bool = False,
                pixel_values: Optional[torch.Tensor] = None, image_grid_thw: Optional[torch.Tensor] = None,
                force_vit_image: bool = False, **unused) -> CausalLMOutput:
        if input_ids is None:
            raise ValueError("You must specify input_ids")
        if pixel_values is not None and image_grid_thw is None:
            raise ValueError("pixel_values requires image_grid_thw")
        # `force_vit_image` (roll/models/model_providers.py:251-320): on a text-only micro-batch the reference runs the
        # ViT on a 4x1176 zero "mock image" and adds mean()*0 to keep ranks in collective lock-step.  The result is
        # numerically a no-op; there is no in-forward collective here, so the flag is accepted and ignored.
        logits = self.engine.forward_logits(input_ids, attention_mask, position_ids, pixel_values, image_grid_thw,
                                            out_dtype=self.logits_dtype)
        return CausalLMOutput(logits=logits, past_key_values=None, rope_deltas=None)

    @torch.no_grad()
    def compute_log_probs(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, position_ids: torch.Tensor,
                          response_mask: torch.Tensor, pixel_values: Optional[torch.Tensor] = None,
                          image_grid_thw: Optional[torch.Tensor] = None, want_entropy: bool = True, **unused):
        """Fused forward of ``ActorWorker.compute_log_probs`` (roll/pipeline/base_worker.py:206-246): (log_probs, entropy)
        fp32 [b, S-1], identical in meaning to ``forward_func_log_probs`` applied to ``forward(...).logits`` -- without
        materialising the logits (Engine.forward_log_probs)."""
        if pixel_values is not None and image_grid_thw is None:
            raise ValueError("pixel_values requires image_grid_thw")
        return self.engine.forward_log_probs(input_ids, attention_mask, position_ids, pixel_values, image_grid_thw, response_mask,
                                             want_entropy=want_entropy)

    @torch.no_grad()
    def get_image_features(self, pixel_values: torch.Tensor, image_grid_thw: torch.Tensor) -> torch.Tensor:
        return self.engine.image_features(pixel_values.to(self.device), image_grid_thw)

    # ---- generate -----------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor = None, attention_mask: Optional[torch.Tensor] = None, use_cache: bool = True,
                 pixel_values: Optional[torch.Tensor] = None, image_grid_thw: Optional[torch.Tensor] = None,
                 position_ids: Optional[torch.Tensor] = None, max_new_tokens: int = 32, do_sample: bool = False,
                 temperature: float = 1.0, top_p: float = 1.0, top_k: int = 0, num_beams: int = 1,
                 repetition_penalty: float = 1.0, num_return_sequences: int = 1,
                 eos_token_id: Union[int, List[int], None] = None, pad_token_id: Optional[int] = None,
                 ignore_eos: bool = False, seed: Optional[int] = None, **unused) -> torch.Tensor:
        """Returns [b * num_return_sequences, S_in + max_new_tokens] like HF generate (prompt included).
        ``do_sample=False`` (or temperature 0 / top_k 1): greedy.  ``do_sample=True``: temperature -> top-k -> top-p sampling
        on the device (the reference's rollout config is T=1, top-p 0.8, top-k 100, examples/infer/rlvr_megatron.yaml:89-95);
        the device sampler keeps at most 1024 candidates: ``top_k`` 0 / None ("disabled" in HF) or > 1024 raises a warning and runs
        with 1024 (with top-p < 1 the tail beyond 1024 candidates carries negligible mass, but it is not the same distribution).
        ``num_return_sequences`` > 1 with sampling draws independent continuations (the prompt batch is repeated, as HF does).
        Decoding stops as soon as every sequence has emitted a stop id (checked on the device, read back every 16 steps)."""
        if num_beams != 1 or repetition_penalty != 1.0:
            raise NotImplementedError("beam search / repetition penalty are not on the SocioSeg path")
        eos = None
        if not ignore_eos and eos_token_id is not None:
            eos = [int(eos_token_id)] if isinstance(eos_token_id, int) else [int(e) for e in eos_token_id]
        pad = int(pad_token_id) if pad_token_id is not None else (eos[0] if eos else 0)
        sample = do_sample and not (temperature == 0 or top_k == 1)
        if sample:
            n = num_return_sequences
            rep = lambda t_: t_.repeat_interleave(n, dim=0) if (t_ is not None and n > 1) else t_
            ids_r, mask_r = rep(input_ids), rep(attention_mask)
            pos_r = position_ids.repeat_interleave(n, dim=1 if position_ids.shape[0] == 3 else 0) if (position_ids is not None and n > 1) else position_ids
            if n > 1 and pixel_values is not None:
                per = (image_grid_thw[:, 0] * image_grid_thw[:, 1] * image_grid_thw[:, 2]).tolist()
                if len(per) != input_ids.shape[0]:
                    raise NotImplementedError("num_return_sequences > 1 needs exactly one image per prompt")
                chunks = torch.split(pixel_values, per, dim=0)
                pixel_values = torch.cat([c for c in chunks for _ in range(n)], 0)
                image_grid_thw = image_grid_thw.repeat_interleave(n, dim=0)
            g = torch.Generator(device=self.device)
            g.manual_seed(int(seed) if seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
            u = torch.rand((ids_r.shape[0], max_new_tokens), generator=g, device=self.device, dtype=torch.float32)
            if top_k in (0, None) or int(top_k) > 1024:
                import warnings
                warnings.warn(f"top_k={top_k} is outside the device sampler's range: sampling from the 1024 most likely tokens", stacklevel=2)
            k = 1024 if top_k in (0, None) else min(int(top_k), 1024)
            new = self.engine.generate_greedy(ids_r, mask_r, pos_r, pixel_values, image_grid_thw, max_new_tokens,
                                              sampling=(float(temperature), k, float(top_p)), uniforms=u, eos_ids=eos, pad_id=pad)
            input_ids, num_return_sequences = ids_r, 1
        else:
            new = self.engine.generate_greedy(input_ids, attention_mask, position_ids, pixel_values, image_grid_thw, max_new_tokens,
                                              eos_ids=eos, pad_id=pad)
        new = new.to(input_ids.device)
        out = torch.cat([input_ids, new], dim=1)
        if num_return_sequences > 1:
            out = out.repeat_interleave(num_return_sequences, dim=0)
        return out
