"""Print (no asserts) every metric test_gpu_golden_3b.py checks, plus host CPU facts for the cpu-baseline thread policy."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
a = torch.randn(4096, 4096); b = torch.randn(4096, 4096)
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt); a @ b
    t = time.time(); a @ b; a @ b; dt = (time.time() - t) / 2
    print(f"threads={nt}: 4096^3 fp32 matmul {dt*1e3:.1f} ms = {2*4096**3/dt/1e12:.2f} TFLOP/s", flush=True)
torch.set_num_threads(32)
from oracle import hf_oracle as O
from socioreasoner_b200 import lib
from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration as Model
def rel_rms(got, ref):
    got, ref = torch.as_tensor(got).float().cpu(), torch.as_tensor(ref).float().cpu()
    return ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
t0 = time.time(); hf = O.build_model("3b"); print("hf build", time.time() - t0, flush=True)
model = Model.from_hf_model(hf, device="cuda"); del hf
b = O.make_batch([0], [448], "3b", n_text=128)
g = np.load("tests/golden/3b.npz")
logits = model(input_ids=b.input_ids, attention_mask=b.attention_mask, position_ids=b.position_ids, use_cache=False,
               pixel_values=b.pixel_values, image_grid_thw=b.image_grid_thw).logits
lv = logits[b.attention_mask.bool()]
print("logits_sub rel_rms", rel_rms(lv[:, ::97], g["logits_sub"]))
print("lse max abs diff", np.abs(torch.logsumexp(lv.float(), -1).cpu().numpy() - g["logits_lse"]).max())
margin = g["logits_top5_val"][:, 0] - g["logits_top5_val"][:, 1]
am = lv.argmax(-1).cpu().numpy()
for thr in (0.02, 0.05, 0.1, 0.2):
    dec = margin > thr
    print(f"top1 agreement margin>{thr}: {(am[dec] == g['logits_top5_idx'][:, 0][dec]).mean():.4f} ({dec.sum()} rows; mism {(am[dec] != g['logits_top5_idx'][:, 0][dec]).sum()})")
B, S, V = logits.shape
labels = b.input_ids[:, 1:].clone()
lp, ent = lib.logprob_entropy(logits[:, :-1].reshape(-1, V), labels.reshape(-1).cuda())
print("logprob max abs diff", np.abs(lp.view(B, S - 1).cpu().numpy() - g["log_probs"]).max(), "entropy max abs diff", np.abs(ent.view(B, S - 1).cpu().numpy() - g["entropy"]).max())
T = g["greedy_tokens"].shape[1]
toks, lg = model.engine.generate_greedy(b.input_ids, b.attention_mask, b.position_ids, b.pixel_values, b.image_grid_thw, T, return_logits=True)
toks_graph = model.engine.generate_greedy(b.input_ids, b.attention_mask, b.position_ids, b.pixel_values, b.image_grid_thw, T)
print("eager", toks.tolist(), "graph", toks_graph.tolist(), "golden", g["greedy_tokens"].tolist())
print("margins", g["greedy_margin"].tolist())
print("decode logits rel_rms per step", [rel_rms(lg[:, t, ::97], g["greedy_logits_sub"][:, t]) for t in range(T)])
