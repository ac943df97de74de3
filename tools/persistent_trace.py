"""In-kernel timeline of the persistent decode step (sr_set_decode_persistent_trace): where do the ~65 us per layer go?
Per phase type (qkv, attention, o-proj, gate/up, down): compute time per CTA, grid-barrier wait, phase wall time, and how far
ahead of the consumers the weight producer runs.  Usage: python tools/persistent_trace.py [layers] [vocab]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import persistent_bringup as PB
from socioreasoner_b200 import lib

n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 36
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 151936
pf_stages = int(sys.argv[3]) if len(sys.argv) > 3 else 0
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib.load().sr_set_decode_persistent_tuning(pf_stages, flags)
print(f"pf_stages={pf_stages} flags={flags}")
torch.manual_seed(0)
layers = PB.make_layers(n_layers)
S = PB.State(8, n_layers, 1024, [900 - 7 * i for i in range(8)], vocab)
P = PB.Persistent(S, layers)
for _ in range(3):
    P.run(0)
torch.cuda.synchronize()
nsm = torch.cuda.get_device_properties(0).multi_processor_count
EV = 10 * n_layers + 8
buf = torch.zeros((nsm, 2, EV, 2), dtype=torch.int64, device="cuda")
lib.load().sr_set_decode_persistent_trace(buf.data_ptr(), EV)
P.run(0)
torch.cuda.synchronize()
lib.load().sr_set_decode_persistent_trace(None, 0)
t = buf.cpu().numpy().astype(np.int64)
gt, ck = t[..., 0], t[..., 1]
prod, cons = 0, 1
t0 = gt[:, cons, 0].min()
total_ns = gt[:, cons, 2 * 5 * n_layers + 1].max() - t0
cyc = (ck[:, cons, 2 * 5 * n_layers + 1] - ck[:, cons, 0]).mean()
ghz = cyc / max(gt[:, cons, 2 * 5 * n_layers + 1].mean() - gt[:, cons, 0].mean(), 1)
print(f"kernel span {total_ns / 1e3:.1f} us, SM clock ~{ghz:.3f} GHz, globaltimer step ~{np.diff(np.unique(gt[0, cons, :50])).min()} ns")
names = ["qkv", "attention", "o-proj", "gate/up", "down"]
lo = 2 if n_layers > 4 else 0
print(f"steady-state layers {lo}..{n_layers - 1}; times in us (clock64 / {ghz:.3f} GHz within a CTA, globaltimer across CTAs)")
print(f"{'phase':10s} {'compute mean':>12s} {'compute max':>12s} {'barrier mean':>12s} {'barrier min':>12s} {'wall':>8s} {'producer lead':>14s}")
tot_wall = 0.0
for k, name in enumerate(names):
    comp, bar, wall, lead = [], [], [], []
    for l in range(lo, n_layers):
        b = 2 * (5 * l + k)          # event index: begin of phase (after previous barrier); b+1 = compute done; b+2 = next begin
        c = (ck[:, cons, b + 1] - ck[:, cons, b]) / ghz / 1e3
        w = (ck[:, cons, b + 2] - ck[:, cons, b + 1]) / ghz / 1e3
        comp.append(c); bar.append(w)
        wall.append((gt[:, cons, b + 2].max() - gt[:, cons, b].max()) / 1e3)
        if k != 1:
            pe = 1 + 4 * l + {0: 0, 2: 1, 3: 2, 4: 3}[k]   # producer event: finished issuing this matrix
            okc = gt[:, prod, pe] > 0   # CTAs without a unit of this matrix record nothing
            lead.append((gt[okc, cons, b].astype(np.float64) - gt[okc, prod, pe]).mean() / 1e3)
    comp, bar = np.array(comp), np.array(bar)
    tot_wall += np.mean(wall)
    print(f"{name:10s} {comp.mean():12.2f} {comp.max(axis=1).mean():12.2f} {bar.mean():12.2f} {bar.min(axis=1).mean():12.2f} {np.mean(wall):8.2f} "
          f"{(np.mean(lead) if lead else float('nan')):14.2f}")
print(f"sum of phase walls per layer: {tot_wall:.2f} us")
b = 2 * 5 * n_layers
head = (gt[:, cons, b + 1].max() - gt[:, cons, b].max()) / 1e3
print(f"LM head: {head:.1f} us ({vocab * 2048 * 2 / head / 1e6:.2f} TB/s)")
# one layer in detail for CTA 0, 1, 147
l = min(5, n_layers - 1)
for cta in (0, 1, nsm // 2, nsm - 1):
    ev = [(gt[cta, cons, 2 * 5 * l + i] - gt[cta, cons, 2 * 5 * l]) / 1e3 for i in range(11)]
    pv = [(gt[cta, prod, 1 + 4 * l + i] - gt[cta, cons, 2 * 5 * l]) / 1e3 for i in range(4)]
    print(f"layer {l} CTA {cta:3d}: consumer events (us from layer start) " + " ".join(f"{x:6.1f}" for x in ev) + " | producer done qkv/o/gu/down " + " ".join(f"{x:6.1f}" for x in pv))
