"""Diagnostic: ragged 3-prompt tiny batch, teacher-forced logits per step vs the HF oracle, for several engine configurations."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hf_oracle as O
from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration as Model
hf = O.build_model("tiny")
model = Model.from_hf_model(hf, device="cuda")
b = O.make_batch([0, 1, 2], [224, 336, 224], "tiny", n_text=[8, 5, 11])
T = 6
toks, lg = O.greedy_decode(hf, b, T)
def rr(a, r): return ((a - r).pow(2).mean(-1).sqrt() / r.pow(2).mean(-1).sqrt())
def run(tag, bb, rows, persistent):
    model.engine.set_persistent_decode(persistent)
    out, l = model.engine.generate_greedy(bb.input_ids, bb.attention_mask, bb.position_ids, bb.pixel_values, bb.image_grid_thw, T,
                                          return_logits=True, forced_tokens=toks[rows])
    l = l.float().cpu()
    for j, r in enumerate(rows):
        e = rr(l[j], lg[r])
        d586 = (l[j][:, 586] - lg[r][:, 586]).tolist(); d331 = (l[j][:, 331] - lg[r][:, 331]).tolist()
        print(f"{tag} sample {r}: rel_rms per step {[round(x, 4) for x in e.tolist()]} argmax {l[j].argmax(-1).tolist()} d586 {[round(x,3) for x in d586]} d331 {[round(x,3) for x in d331]}")
for p in (True, False):
    run(f"batch3 persistent={p}", b, [0, 1, 2], p)
    b2 = O.make_batch([2], [224], "tiny", n_text=[11])
    run(f"alone-unpadded persistent={p}", b2, [2], p)
    import copy
    b3 = O.Batch(b.input_ids[2:3], b.attention_mask[2:3], b.position_ids[:, 2:3], b.pixel_values[-256:], b.image_grid_thw[2:3])
    run(f"alone-leftpadded persistent={p}", b3, [2], p)
