"""Generate the committed golden vectors under tests/golden/ from the HF-eager CPU oracle.

    python tools/make_golden.py tiny        # seconds
    python tools/make_golden.py 3b          # ~5 min on 8 cores (3.75 B parameter fp32 model)

The reference ships no golden vectors, so parity is pinned on the third-party implementation the
reference executes (HF transformers Qwen2.5-VL eager, driven like roll/distributed/strategy/hf_strategy.py:84-90) with
seeded random-init weights.  Inputs are regenerated from seeds by the tests; the fixture stores input checksums (to pin
the input pipeline) and compact output summaries (to pin the arithmetic).
"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hf_oracle as O, host_ref

CASES = {
    # name: (config, tile ids, sides, n_text, decode steps)
    "tiny": ("tiny", [0, 1], [224, 336], 8, 6),
    "3b": ("3b", [0], [448], 128, 8),
}


def summarize(hf, b, steps):
    out = {}
    out["in_pixel_sum"] = np.float64(b.pixel_values.double().sum().item())
    out["in_pixel_abs_sum"] = np.float64(b.pixel_values.double().abs().sum().item())
    out["in_ids_sum"] = np.int64(b.input_ids.sum().item())
    out["in_pos_sum"] = np.int64(b.position_ids.sum().item())
    out["grid"] = b.image_grid_thw.numpy()
    merged, hid, caps = O.vit_forward(hf, b.pixel_values, b.image_grid_thw, capture_blocks=True)
    out["vit_merged_sub"] = merged[::7, :64].numpy().astype(np.float32)
    out["vit_merged_rms"] = np.float64(merged.double().pow(2).mean().sqrt().item())
    out["vit_block_rms"] = np.array([c.double().pow(2).mean().sqrt().item() for c in caps])
    out["vit_last_hidden_sub"] = hid[::37, :32].numpy().astype(np.float32)
    res = O.forward_logits(hf, b)
    logits = res.logits.float()
    keep = b.attention_mask.bool()
    lv = logits[keep]
    top = lv.topk(5, -1)
    out["logits_top5_val"] = top.values.numpy()
    out["logits_top5_idx"] = top.indices.numpy()
    out["logits_lse"] = torch.logsumexp(lv, -1).numpy()
    out["logits_sub"] = lv[:, ::97].numpy()
    out["log_probs"] = host_ref.op_compute_log_probs(logits, b.input_ids, b.attention_mask).numpy()
    out["entropy"] = host_ref.op_compute_entropy(logits, b.attention_mask).numpy()
    toks, lg = O.greedy_decode(hf, b, steps)
    out["greedy_tokens"] = toks.numpy()
    t2 = lg.topk(2, -1).values
    out["greedy_margin"] = (t2[..., 0] - t2[..., 1]).numpy()
    out["greedy_logits_sub"] = lg[:, :, ::97].numpy()
    return out


if __name__ == "__main__":
    name = sys.argv[1]
    cfg, ids, sides, n_text, steps = CASES[name]
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    hf = O.build_model(cfg)
    print("built", time.time() - t0, flush=True)
    b = O.make_batch(ids, sides, cfg, n_text=n_text)
    out = summarize(hf, b, steps)
    out["meta"] = np.array([f"case={name} cfg={cfg} tiles={ids} sides={sides} n_text={n_text} steps={steps} torch={torch.__version__} "
                            f"transformers={__import__('transformers').__version__} seed=42 perturb=True round_bf16=True"])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes in", time.time() - t0, "s")
