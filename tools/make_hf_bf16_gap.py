"""HF-eager bf16 (CPU) vs the fp32 golden of the 3B case: the oracle's own bf16-vs-fp32 gap -> tests/golden/3b_hf_bf16_gap.json
(values copied into the json by hand from this script's printout; ~5 min on 8 cores)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from oracle import hf_oracle as O, host_ref
torch.set_num_threads(8)
g = np.load("/root/repo/tests/golden/3b.npz")
t0 = time.time()
hf = O.build_model("3b", dtype=torch.bfloat16)
print("built", time.time() - t0, flush=True)
b = O.make_batch([0], [448], "3b", n_text=128)
out = O.forward_logits(hf, b).logits.float()
print("fwd", time.time() - t0, flush=True)
keep = b.attention_mask.bool()
lv = out[keep]
ref = torch.from_numpy(g["logits_sub"])
got = lv[:, ::97]
rr = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
lp = host_ref.op_compute_log_probs(out, b.input_ids, b.attention_mask).numpy()
ent = host_ref.op_compute_entropy(out, b.attention_mask).numpy()
print("HF bf16 vs fp32: logits_sub rel_rms", rr, "logprob maxdiff", np.abs(lp - g["log_probs"]).max(), "entropy maxdiff", np.abs(ent - g["entropy"]).max(),
      "lse maxdiff", np.abs(torch.logsumexp(lv, -1).numpy() - g["logits_lse"]).max())
margin = g["logits_top5_val"][:, 0] - g["logits_top5_val"][:, 1]
am = lv.argmax(-1).numpy()
for thr in (0.05, 0.1):
    dec = margin > thr
    print("top1 agree margin>", thr, (am[dec] == g["logits_top5_idx"][:, 0][dec]).mean())
np.savez("/tmp/hf16_gap.npz", logits_rel_rms=rr, logprob_maxdiff=np.abs(lp - g["log_probs"]).max(), entropy_maxdiff=np.abs(ent - g["entropy"]).max())
