"""Test helper: a complete local HF checkpoint directory (config.json + safetensors + tokenizer + image processor) for the tiny
oracle configuration -- what ``model_args.model_name_or_path`` points at in the reference (models are hub names there; here
there is no network, so the files are written from the seeded random-init oracle)."""
import json
import os


def tiny_tokenizer(vocab_size: int = 8192):
    """A fast tokenizer over a WordLevel vocabulary "t0" .. "t<V-1>" with the Qwen special-token NAMES at the tiny config's ids."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    special = {7001: "<|im_start|>", 7002: "<|im_end|>", 7000: "<|endoftext|>", 8001: "<|image_pad|>", 8002: "<|video_pad|>",
               8003: "<|vision_start|>", 8004: "<|vision_end|>"}
    vocab = {special.get(i, f"t{i}"): i for i in range(vocab_size)}
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="t0"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    return PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|im_end|>", pad_token="<|endoftext|>", padding_side="left",
                                   additional_special_tokens=["<|im_start|>", "<|image_pad|>", "<|vision_start|>", "<|vision_end|>"])


def write_checkpoint(path: str, hf_model, min_pixels: int = 56 * 56, max_pixels: int = 768 * 768) -> str:
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor

    os.makedirs(path, exist_ok=True)
    hf_model.save_pretrained(path, safe_serialization=True)
    tiny_tokenizer(hf_model.config.text_config.vocab_size).save_pretrained(path)
    Qwen2VLImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels).save_pretrained(path)
    assert os.path.exists(os.path.join(path, "config.json"))
    return path
