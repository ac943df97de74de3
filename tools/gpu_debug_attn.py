import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib, plan
nimg = int(sys.argv[1]); kind = sys.argv[2]; H = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = "cuda"
vp = plan.vit_plan(np.array([[1, 54, 54]] * nimg))
pl = vp.window_plan if kind == "win" else vp.full_plan
n = vp.n
torch.manual_seed(0)
q = torch.zeros(n, H, 128, device=dev, dtype=torch.bfloat16); k = torch.zeros_like(q)
q[:, :, :80] = torch.randn(n, H, 80, device=dev).bfloat16(); k[:, :, :80] = torch.randn(n, H, 80, device=dev).bfloat16()
npad = (n + 7) // 8 * 8
vt = torch.zeros(H * 80, npad, device=dev, dtype=torch.bfloat16)
vt[:, :n] = torch.randn(H * 80, n, device=dev).bfloat16()
out = torch.zeros(n, H, 80, device=dev, dtype=torch.bfloat16)
tiles = torch.from_numpy(pl.tiles).to(dev); lo = torch.from_numpy(pl.row_lo).to(dev); hi = torch.from_numpy(pl.row_hi).to(dev)
lib.attention(q, k, vt, out, tiles, lo, hi, 80 ** -0.5, 80)
torch.cuda.synchronize()
print(f"OK nimg={nimg} kind={kind} H={H} n={n} tiles={len(pl.tiles)} out_mean={out.float().abs().mean().item():.4f}", flush=True)
