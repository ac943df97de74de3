"""``roll/models/model_providers.py:323-372`` surface for the actor model."""
from __future__ import annotations


def default_actor_model_provider(tokenizer, model_args, training_args=None, is_trainable: bool = False):
    """Same signature as the reference (model_providers.py:323).  Returns the B200 module whose ``forward(...).logits``
    and ``generate(...)`` the strategies call.  ``model_args.model_name_or_path`` must be an HF checkpoint directory
    (config.json + safetensors; transformers-4.x or 5.x key layout)."""
    if is_trainable:
        raise NotImplementedError("the B200 path is inference-only (rollout / log-prob forward); training is out of scope")
    from ..modeling import SocioReasonerB200ForConditionalGeneration

    path = getattr(model_args, "model_name_or_path", None) or model_args
    return SocioReasonerB200ForConditionalGeneration.from_pretrained(path)
