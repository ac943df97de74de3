"""In-kernel timeline of the tensor-memory attention kernel (variant 2) at the bench geometry: clock64 stamps of the first work items
of CTAs 0..3 (sr_set_attention_trace).  Prints, per case, the mean duration of each segment of an item in microseconds.
Usage: python tools/attn_trace.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib, plan

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
GHZ = 1.9  # clock64 runs at the SM clock (reported max 1.965 GHz; the kernel is short, assume ~1.9)
SOFT = ["wait S", "load S (TMEM)", "max (+mask)", "exp + P store issue", "st wait + arrive", "wait P.V (last block)", "epilogue (O load, scale, store)"]


def case(name, pl, n_rows, n_keys, H, Hkv, hdv, d_real):
    q = torch.zeros(n_rows, H, 128, dtype=torch.bfloat16, device=dev); q[:, :, :d_real] = rn(n_rows, H, d_real)
    k = torch.zeros(n_keys, Hkv, 128, dtype=torch.bfloat16, device=dev); k[:, :, :d_real] = rn(n_keys, Hkv, d_real)
    ncol = (n_keys + 7) // 8 * 8
    vt = torch.zeros(Hkv * hdv, ncol, dtype=torch.bfloat16, device=dev); vt[:, :n_keys] = rn(Hkv * hdv, n_keys)
    tiles, lo, hi = torch.from_numpy(pl.tiles).to(dev), torch.from_numpy(pl.row_lo).to(dev), torch.from_numpy(pl.row_hi).to(dev)
    out = torch.zeros(n_rows, H, hdv, dtype=torch.bfloat16, device=dev)
    lib.load().sr_set_attention_lean(2)
    run = lambda: lib.attention(q, k, vt, out, tiles, lo, hi, d_real ** -0.5, hdv)
    for _ in range(3):
        run()
    buf = torch.zeros(4 * 16 * 16, dtype=torch.int64, device=dev)
    lib.load().sr_set_attention_trace(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.load().sr_set_attention_trace(None)
    t = buf.cpu().numpy().reshape(4, 16, 16).astype(np.float64)
    print(f"--- {name}: {tiles.shape[0]} tiles x {H} heads, blocks per item {int((pl.tiles[0, 3] - pl.tiles[0, 2] + 127) // 128)}")
    for cta in range(2):
        items = [i for i in range(1, 16) if t[cta, i, 7] > 0 and t[cta, i, 0] > 0]
        if not items:
            continue
        seg = np.array([[t[cta, i, s + 1] - t[cta, i, s] for s in range(7)] for i in items]) / GHZ / 1e3
        period = np.diff([t[cta, i, 0] for i in items]) / GHZ / 1e3
        print(f"CTA {cta}: item period {period.mean():.2f} us (softmax thread, first block of the item):")
        for n_, v in zip(SOFT, seg.mean(0)):
            print(f"    {n_:34s} {v:7.2f} us")
        mm = np.array([[t[cta, i, 9] - t[cta, i, 8], t[cta, i, 10] - t[cta, i, 9], t[cta, i, 11] - t[cta, i, 10]] for i in items]) / GHZ / 1e3
        print(f"    MMA thread: wait Q/K {mm[:, 0].mean():.2f} us, issue S + wait V {mm[:, 1].mean():.2f} us, wait P {mm[:, 2].mean():.2f} us")


vp = plan.vit_plan(np.array([[1, 54, 54]] * 8))
N = vp.n
case("ViT windowed", vp.window_plan, N, N, 16, 16, 80, 80)
case("ViT full", vp.full_plan, N, N, 16, 16, 80, 80)
