"""A/B timing of the persistent decode step at the bench geometry: one configuration per argument
`flags[:ring_stages[:q,o,g,d unit offsets]]`, all on the same weights.  SR_LIB_PATH selects an alternative build of the same C ABI.
Usage: python tools/persistent_ab.py 1 0 1:10 1:6 1:8:96,40,0,20 ..."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import persistent_bringup as PB
from socioreasoner_b200 import lib

n_layers, vocab = 36, 151936
def parse(a):
    parts = a.split(":")
    flags = int(parts[0])
    ring = int(parts[1]) if len(parts) > 1 and parts[1] else 0
    offs = tuple(int(v) for v in parts[2].split(",")) if len(parts) > 2 else None
    return flags, ring, offs


cfgs = [parse(a) for a in sys.argv[1:]] or [(17, 0, None)]
torch.manual_seed(0)
layers = PB.make_layers(n_layers)
S = PB.State(8, n_layers, 1024, [900 - 7 * i for i in range(8)], vocab)
P = PB.Persistent(S, layers)
wbytes = n_layers * (PB.QKV * PB.H + PB.H * PB.H + 2 * PB.I * PB.H + PB.H * PB.I) * 2 + vocab * PB.H * 2
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
print("lib:", lib.LIB_PATH)
for flags, ring, offs in cfgs:
    lib.decode_persistent_tuning(flags=flags, ring_stages=ring, offsets=offs)
    P.reset()
    for _ in range(3):
        P.run(0)
    torch.cuda.synchronize()
    P.reset(); P.run(0); torch.cuda.synchronize()
    if ref is None:
        ref = P.logits.clone()
    same = torch.equal(ref, P.logits)
    e0.record()
    for _ in range(20):
        P.run(0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20):
        P.run(5 * n_layers)
    e1.record(); torch.cuda.synchronize()
    ms_l = e0.elapsed_time(e1) / 20
    print(f"flags={flags:3d} ring={ring:2d} offsets={offs}: {ms:.3f} ms/step ({wbytes / ms / 1e9:.2f} TB/s), {ms_l / n_layers * 1e3:.1f} us/layer, "
          f"LM head {1e3 * (ms - ms_l):.0f} us, same logits: {same}", flush=True)
