"""Bring-up parity run: B200 engine vs the HF-eager CPU oracle on a reduced-depth config (prints statistics)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hf_oracle as O
from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration as M

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
sides = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [448, 224]
torch.set_num_threads(os.cpu_count())

def rel(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    e = (got - ref).abs()
    print(f"{name}: max_abs={e.max():.4e} mean_abs={e.mean():.4e} ref_absmax={ref.abs().max():.3e} ref_rms={ref.pow(2).mean().sqrt():.3e} "
          f"rel_rms={(e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.3e} nan={torch.isnan(got).sum().item()}", flush=True)

t0 = time.time()
hf = O.build_model(name)
print("oracle built", time.time() - t0, flush=True)
model = M.from_hf_model(hf)
b = O.make_batch(list(range(len(sides))), sides, name, n_text=24)
print("batch", b.input_ids.shape, b.pixel_values.shape, b.image_grid_thw.tolist(), flush=True)

# oracle
t0 = time.time()
merged_ref, hid_ref, caps_ref = O.vit_forward(hf, b.pixel_values, b.image_grid_thw, capture_blocks=True)
out_ref = O.forward_logits(hf, b, output_hidden_states=True)
print("oracle fwd", time.time() - t0, flush=True)
hf16 = hf.to(torch.bfloat16)
out16 = O.forward_logits(hf16, b)
merged16, _, caps16 = O.vit_forward(hf16, b.pixel_values, b.image_grid_thw, capture_blocks=True)
hf = hf.to(torch.float32)

# engine
caps = []
merged, vp = model.engine.vit(b.pixel_values.cuda(), b.image_grid_thw, capture=caps)
torch.cuda.synchronize()
for i, (c, r) in enumerate(zip(caps, caps_ref)):
    rel(f"vit block {i} hidden", c, r)
    rel(f"   (HF bf16 CPU vs fp32)", caps16[i], r)
feat = model.get_image_features(b.pixel_values, b.image_grid_thw)
rel("vit merged (orig order)", feat, merged_ref)
rel("   (HF bf16 CPU vs fp32)", merged16, merged_ref)

dcaps = []
x, sp, cache, pos3 = model.engine.prefill(b.input_ids, b.attention_mask, b.position_ids, b.pixel_values, b.image_grid_thw, capture=dcaps)
keep = b.attention_mask.reshape(-1) != 0
for i, c in enumerate(dcaps):
    ref = out_ref.hidden_states[i + 1].reshape(-1, c.shape[-1])[keep]
    if i == len(dcaps) - 1:
        break  # HF's last hidden_state is post-norm
    rel(f"decoder layer {i} hidden", c, ref)
logits = model(input_ids=b.input_ids, attention_mask=b.attention_mask, position_ids=b.position_ids, use_cache=False,
               pixel_values=b.pixel_values, image_grid_thw=b.image_grid_thw, force_vit_image=True).logits
V = logits.shape[-1]
rel("logits (valid rows)", logits.reshape(-1, V)[keep], out_ref.logits.reshape(-1, V)[keep])
rel("   (HF bf16 CPU vs fp32)", out16.logits.reshape(-1, V)[keep], out_ref.logits.reshape(-1, V)[keep])
am = logits.reshape(-1, V)[keep].argmax(-1).cpu()
amr = out_ref.logits.reshape(-1, V)[keep].argmax(-1)
print("argmax agreement", (am == amr).float().mean().item(), " HF-bf16:", (out16.logits.reshape(-1, V)[keep].float().argmax(-1) == amr).float().mean().item())

T = 8
toks_ref, lg_ref = O.greedy_decode(hf, b, T)
toks, lg = model.engine.generate_greedy(b.input_ids, b.attention_mask, b.position_ids, b.pixel_values, b.image_grid_thw, T, return_logits=True)
print("greedy ref", toks_ref.tolist())
print("greedy got", toks.cpu().tolist())
rel("decode logits (all steps)", lg, lg_ref)
# teacher-forced check independent of argmax ties: margin of oracle top-1 vs top-2
top2 = lg_ref.topk(2, -1).values
print("oracle top1-top2 margins", (top2[..., 0] - top2[..., 1]).flatten().tolist())
