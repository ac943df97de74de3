"""In-kernel timeline of the persistent decode step (sr_set_decode_persistent_trace): where does a layer's time go?
The kernel has no grid barrier; its consumers record 9 events per layer (CTA-local thread 32):
  0 A-staged (x polled + normed)   1 A-done (qkv written)       2 B-q-ready (q/k/v polled)   3 B-partial-published
  4 C-staged (attention out polled) 5 C-done (x after o-proj)    6 D-staged (x polled + normed) 7 DE-done (gate/up + down slices
  streamed, fp32 partial of x published)   8 E-done (partials of all CTAs reduced, x written)
Reported per segment: mean / max over CTAs of the segment's duration (us) in the steady-state layers, i.e. the wait for the
other CTAs' data (the "staged" / "ready" segments) versus the streaming compute (the "done" segments), plus the producer's lead.
Usage: python tools/persistent_trace.py [layers] [vocab] [flags] [ring]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import persistent_bringup as PB
from socioreasoner_b200 import lib

n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 36
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 151936
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ring = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib.decode_persistent_tuning(flags=flags, ring_stages=ring)
print(f"flags={flags} ring={ring or 'default'}")
torch.manual_seed(0)
layers = PB.make_layers(n_layers)
S = PB.State(8, n_layers, 1024, [900 - 7 * i for i in range(8)], vocab)
P = PB.Persistent(S, layers)
for _ in range(3):
    P.run(0)
torch.cuda.synchronize()
nsm = torch.cuda.get_device_properties(0).multi_processor_count
EPL = 9
EV = EPL * n_layers + 8
buf = torch.zeros((nsm, 2, EV, 2), dtype=torch.int64, device="cuda")
lib.load().sr_set_decode_persistent_trace(buf.data_ptr(), EV)
P.run(0)
torch.cuda.synchronize()
lib.load().sr_set_decode_persistent_trace(None, 0)
t = buf.cpu().numpy().astype(np.int64)
if os.environ.get("SR_TRACE_DUMP"):
    np.save(os.environ["SR_TRACE_DUMP"], t)
gt, ck = t[..., 0], t[..., 1]
prod, cons = 0, 1
last = 1 + EPL * n_layers          # LM head done
t0 = gt[:, cons, 0].min()
total_ns = gt[:, cons, last].max() - t0
cyc = (ck[:, cons, last] - ck[:, cons, 0]).mean()
ghz = cyc / max(gt[:, cons, last].mean() - gt[:, cons, 0].mean(), 1)
print(f"kernel span {total_ns / 1e3:.1f} us, SM clock ~{ghz:.3f} GHz")
names = ["A wait x", "A qkv gemv", "B wait q", "B attention", "C wait ao(+merge)", "C o-proj", "D wait x", "DE gate/up+down", "E reduce partials"]
lo = 2 if n_layers > 4 else 0
print(f"steady-state layers {lo}..{n_layers - 1}; per segment: duration in us (clock64 / {ghz:.3f} GHz within a CTA)")
print(f"{'segment':20s} {'mean':>8s} {'p50':>8s} {'max':>8s}")
tot = 0.0
for k, name in enumerate(names):
    d = []
    for l in range(lo, n_layers):
        e1 = 1 + EPL * l + k
        e0 = e1 - 1 if (k > 0 or l > 0) else 0
        d.append((ck[:, cons, e1] - ck[:, cons, e0]) / ghz / 1e3)
    d = np.array(d)
    tot += d.mean()
    print(f"{name:20s} {d.mean():8.2f} {np.median(d):8.2f} {d.max(axis=1).mean():8.2f}")
lay = [(gt[:, cons, 1 + EPL * (l + 1) - 1].max() - gt[:, cons, 1 + EPL * l - 1].max()) / 1e3 for l in range(lo, n_layers - 1)]
print(f"sum of segment means {tot:.2f} us; layer period (max over CTAs of E-done, layer to layer) {np.mean(lay):.2f} us")
head = (gt[:, cons, last].max() - gt[:, cons, last - 1].max()) / 1e3
print(f"LM head: {head:.1f} us ({vocab * 2048 * 2 / head / 1e6:.2f} TB/s)")
for k, nm in ((0, "qkv"), (1, "o"), (2, "gate/up"), (3, "down")):
    lead = []
    cons_ev = {0: 1, 1: 5, 2: 7, 3: 8}[k]   # consumer finished the matrix
    for l in range(lo, n_layers):
        pe = 1 + 4 * l + k
        okc = gt[:, prod, pe] > 0
        lead.append((gt[okc, cons, 1 + EPL * l + cons_ev].astype(np.float64) - gt[okc, prod, pe]).mean() / 1e3)
    print(f"producer finished issuing {nm:8s} {np.mean(lead):7.2f} us before its consumers finished it")
l = min(5, n_layers - 1)
for cta in (0, 1, nsm // 2, nsm - 1):
    base = gt[cta, cons, 1 + EPL * l - 1]
    ev = [(gt[cta, cons, 1 + EPL * l + i] - base) / 1e3 for i in range(EPL)]
    pv = [(gt[cta, prod, 1 + 4 * l + i] - base) / 1e3 for i in range(4)]
    print(f"layer {l} CTA {cta:3d}: events (us from previous E-done) " + " ".join(f"{x:6.1f}" for x in ev) + " | producer done qkv/o/gu/down " + " ".join(f"{x:6.1f}" for x in pv))

# absolute timeline (globaltimer): events relative to the moment the LAST CTA finished the previous layer's down projection
evn = ["A-staged", "A-done", "B-qready", "B-published", "C-staged", "C-done", "D-staged", "D-done", "E-done"]
rows = []
for l in range(max(lo, 1), n_layers - 1):
    base = gt[:, cons, 1 + EPL * l - 1].max()
    rows.append(gt[:, cons, 1 + EPL * l:1 + EPL * (l + 1)] - base)
ev = np.mean(rows, axis=0) / 1e3
print("absolute timeline, us after the last CTA finished the previous layer (mean over layers; min / p10 / p50 / p90 / max over CTAs)")
for k, nme in enumerate(evn):
    v = ev[:, k]
    print(f"  {nme:12s} {v.min():7.2f} {np.percentile(v, 10):7.2f} {np.median(v):7.2f} {np.percentile(v, 90):7.2f} {v.max():7.2f}")
