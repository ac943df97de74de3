"""``roll/models/model_providers.py`` surface the strategies' ``initialize`` goes through: tokenizer / processor / actor model
providers with the reference's signatures (``default_tokenizer_provider`` :36-45, ``default_processor_provider`` :48-55,
``default_actor_model_provider`` :323-372).  ``download_model`` of the reference only resolves hub names; paths here are local
checkpoint directories (there is no network on the box)."""
from __future__ import annotations


def _path(model_args):
    return getattr(model_args, "model_name_or_path", None) or model_args


def default_tokenizer_provider(model_args):
    """model_providers.py:36-45: fast tokenizer, left padding (the collator left-pads prompts, collator.py:421-565)."""
    from transformers import AutoTokenizer

    return AutoTokenizer.from_pretrained(_path(model_args), use_fast=True, split_special_tokens=False, trust_remote_code=True,
                                         padding_side="left")


def default_processor_provider(model_args):
    """model_providers.py:48-55: the processor is optional (text-only checkpoints have none) -> None instead of raising."""
    from transformers import AutoProcessor

    try:
        return AutoProcessor.from_pretrained(_path(model_args), trust_remote_code=True)
    except Exception:  # noqa: BLE001  (same behaviour as the reference: log-and-continue)
        return None


def default_actor_model_provider(tokenizer, model_args, training_args=None, is_trainable: bool = False):
    """Same signature as the reference (model_providers.py:323).  Returns the B200 module whose ``forward(...).logits``
    and ``generate(...)`` the strategies call.  ``model_args.model_name_or_path`` must be an HF checkpoint directory
    (config.json + safetensors; transformers-4.x or 5.x key layout)."""
    if is_trainable:
        raise NotImplementedError("the B200 path is inference-only (rollout / log-prob forward); training is out of scope")
    from ..modeling import SocioReasonerB200ForConditionalGeneration

    return SocioReasonerB200ForConditionalGeneration.from_pretrained(_path(model_args))
