"""tcgen05 attention kernel at the bench geometry (8 tiles of 54x54 patches; 8 prompts of 867 tokens): the one-CTA-per-SM
variant against the lean two-CTAs-per-SM variants (P in shared memory / P in tensor memory) -- bit-identical outputs, CUDA-event timings per attention shape.
Usage: python tools/attn_bench.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib, plan

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()


def case(name, pl, n_rows, n_keys, H, Hkv, hdv, d_real, flops):
    q = torch.zeros(n_rows, H, 128, dtype=torch.bfloat16, device=dev); q[:, :, :d_real] = rn(n_rows, H, d_real)
    k = torch.zeros(n_keys, Hkv, 128, dtype=torch.bfloat16, device=dev); k[:, :, :d_real] = rn(n_keys, Hkv, d_real)
    ncol = (n_keys + 7) // 8 * 8
    vt = torch.zeros(Hkv * hdv, ncol, dtype=torch.bfloat16, device=dev); vt[:, :n_keys] = rn(Hkv * hdv, n_keys)
    tiles, lo, hi = torch.from_numpy(pl.tiles).to(dev), torch.from_numpy(pl.row_lo).to(dev), torch.from_numpy(pl.row_hi).to(dev)
    outs, times = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for lean in (0, 1, 3, 2):   # 3 = variant 2 with one CTA per item; 2 (last, left selected) = the product configuration
        lib.load().sr_set_attention_lean(lean)
        out = torch.zeros(n_rows, H, hdv, dtype=torch.bfloat16, device=dev)
        run = lambda: lib.attention(q, k, vt, out, tiles, lo, hi, d_real ** -0.5, hdv)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        outs.append(out.clone()); times.append(e0.elapsed_time(e1) / 10 * 1e3)
    # variants 0 and 1 are bit-identical; variant 2 keeps a lazy softmax reference point (same mathematics, other rounding)
    rel = ((outs[0].float() - outs[3].float()).norm() / outs[0].float().norm()).item()
    same = torch.equal(outs[0], outs[1]) and torch.equal(outs[2], outs[3]) and rel < 4e-3
    finite = torch.isfinite(outs[3].float()).all().item()
    maxdiff = (outs[0].float() - outs[3].float()).abs().max().item()
    # same inputs with Q / K stored head-major [H, rows, 128] (a tile's Q / K of one head = one contiguous 32 KB block)
    lib.load().sr_set_attention_lean(2)
    qh = q.permute(1, 0, 2).contiguous().permute(1, 0, 2)
    kh = k.permute(1, 0, 2).contiguous().permute(1, 0, 2)
    out_h = torch.zeros(n_rows, H, hdv, dtype=torch.bfloat16, device=dev)
    run_h = lambda: lib.attention(qh, kh, vt, out_h, tiles, lo, hi, d_real ** -0.5, hdv)
    for _ in range(3):
        run_h()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run_h()
    e1.record(); torch.cuda.synchronize()
    t_h = e0.elapsed_time(e1) / 10 * 1e3
    same_h = torch.equal(out_h, outs[3])
    print(f"{name:28s} head-major Q/K: {t_h:8.1f} us (same bits: {same_h})")
    print(f"{name:28s} tiles={tiles.shape[0]:5d} x {H:2d} heads  1 CTA/SM {times[0]:8.1f} us ({flops / times[0] / 1e6:6.1f} TF/s)   "
          f"lean {times[1]:8.1f} us ({flops / times[1] / 1e6:6.1f} TF/s)   P,O in TMEM {times[2]:8.1f} us, persistent grid {times[3]:8.1f} us ({flops / times[3] / 1e6:6.1f} TF/s)   "
          f"0 == 1 and 2 == 3 bitwise, 2 vs 0 within 4e-3: {same} (rel {rel:.2e}, max diff {maxdiff:.3g})  finite: {finite}", flush=True)
    return same and finite


ok = True
vp = plan.vit_plan(np.array([[1, 54, 54]] * 8))
N = vp.n
wflops = 4.0 * 80 * 16 * float(((vp.window_plan.row_hi - vp.window_plan.row_lo).astype(np.float64)).sum())
ok &= case("ViT windowed (28 layers)", vp.window_plan, N, N, 16, 16, 80, 80, wflops)
ok &= case("ViT full (4 layers)", vp.full_plan, N, N, 16, 16, 80, 80, 4.0 * 80 * 16 * 8 * 2916.0 ** 2)
lens, L = [867] * 8, 1024
qs = np.concatenate([[0], np.cumsum(lens)[:-1]]); ks = np.arange(8) * L
cp = plan.plan_causal(lens, qs, ks)
ok &= case("decoder causal prefill (36)", cp, sum(lens), 8 * L, 16, 2, 128, 128, 4.0 * 128 * 16 * 8 * 867.0 ** 2 / 2)
print("ALL OK" if ok else "MISMATCH")
