"""Teacher-forced decode golden at the BENCH geometry -> tests/golden/3b_decode.npz   (~10 min on 8 cores)

    python tools/make_golden_decode.py

Two prompts through the HF-eager fp32 oracle (oracle/hf_oracle.py, driven like roll/distributed/strategy/hf_strategy.py:84-90,
decode positions per roll/utils/functionals.py:838-852), 3B dimensions, seeded weights:
  A  one 1024^2 tile -> 756^2 (N = 2916 patches, 729 vision tokens), 128 text ids: S = 867 -- BASELINE.json configs[1], the
     geometry bench.py measures;
  B  TWO 756^2 tiles in one prompt (the reference's map + sat prompt, rlvr_socioseg_vlm_pipeline_infer.py:74-81): S = 1598,
     beyond the persistent decode kernel's single-pass range at batch 8 (1152 keys).
For each: T = 32 greedy steps.  The oracle feeds its OWN tokens back, and the GPU test feeds the oracle's tokens too (teacher
forcing), so every one of the 32 logit rows is comparable regardless of near-tie argmax flips.  Stored per step: the token, the
logits sub-sampled ::97, top-5 values / indices, logsumexp, top-1 margin.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hf_oracle as O

CASES = {
    # name: (tile ids, sides, images_per_prompt, n_text)
    "A": ([20], [1024], 1, 128),
    "B": ([21, 22], [1024, 1024], 2, 128),
}
STEPS = 32


def case_batch(name):
    ids, sides, k, n_text = CASES[name]
    return O.make_batch(ids, sides, "3b", n_text=n_text, images_per_prompt=k)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    hf = O.build_model("3b")
    print("built", time.time() - t0, flush=True)
    out = {}
    for name in CASES:
        b = case_batch(name)
        toks, lg = O.greedy_decode(hf, b, STEPS)
        lg = lg[0].float()
        top = lg.topk(5, -1)
        out[f"{name}_in_ids_sum"] = np.int64(b.input_ids.sum().item())
        out[f"{name}_in_pos_sum"] = np.int64(b.position_ids.sum().item())
        out[f"{name}_in_pixel_sum"] = np.float64(b.pixel_values.double().sum().item())
        out[f"{name}_prompt_len"] = np.int64(b.input_ids.shape[1])
        out[f"{name}_tokens"] = toks[0].numpy()
        out[f"{name}_logits_sub"] = lg[:, ::97].numpy()
        out[f"{name}_top5_val"] = top.values.numpy()
        out[f"{name}_top5_idx"] = top.indices.numpy()
        out[f"{name}_lse"] = torch.logsumexp(lg, -1).numpy()
        print(name, "S", b.input_ids.shape[1], "tokens", toks[0].tolist(), "margins", (top.values[:, 0] - top.values[:, 1]).tolist(),
              time.time() - t0, flush=True)
    out["meta"] = np.array([f"cases={CASES} steps={STEPS} torch={torch.__version__} transformers={__import__('transformers').__version__} "
                            f"seed=42 perturb=True round_bf16=True fp32 eager"])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "3b_decode.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes in", time.time() - t0, "s")
