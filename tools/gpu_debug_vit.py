import os, sys, time, traceback
os.environ["SR_SYNC_DEBUG"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib
from socioreasoner_b200.config import SocioReasonerConfig
from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration as M
cfg = SocioReasonerConfig()
model = M.random_init(cfg, seed=42)
print("model built", flush=True)
ntile = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grid = np.array([[1, 54, 54]] * ntile)
pv = torch.randn(ntile * 2916, 1176, device="cuda")
orig_gemm, orig_att = lib.gemm, lib.attention
cnt = [0]
def wrap(fn, name):
    def f(*a, **k):
        cnt[0] += 1
        try:
            return fn(*a, **k)
        except Exception:
            print(f"FAILED at call #{cnt[0]} {name} shapes", [tuple(x.shape) for x in a if torch.is_tensor(x)], {kk: (tuple(v.shape) if torch.is_tensor(v) else v) for kk, v in k.items()}, flush=True)
            raise
    return f
for n in ("gemm", "attention", "rmsnorm", "gather_rows", "vit_rope_split", "vit_rope_table"):
    setattr(lib, n, wrap(getattr(lib, n), n))
try:
    caps = []
    merged, vp = model.engine.vit(pv, grid, capture=caps)
    torch.cuda.synchronize()
    print("vit ok", merged.float().abs().mean().item(), len(caps), "calls", cnt[0])
except Exception:
    traceback.print_exc()
