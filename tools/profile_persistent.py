"""One persistent decode step at the bench geometry (36 layers, vocab 151936, B=8, ~900 cached tokens) for ncu:
ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 3 -c 1 python tools/profile_persistent.py [flags] [ring]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import persistent_bringup as PB
from socioreasoner_b200 import lib
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ring = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib.decode_persistent_tuning(flags=flags, ring_stages=ring)
torch.manual_seed(0)
layers = PB.make_layers(36)
S = PB.State(8, 36, 1024, [900 - 7 * i for i in range(8)], 151936)
P = PB.Persistent(S, layers)
for _ in range(4):
    P.run(0)
torch.cuda.synchronize()
print("done")
