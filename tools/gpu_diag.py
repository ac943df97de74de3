"""Bring-up diagnostics: run every kernel of libsocioreasoner_b200.so against torch (GPU) references and print error
statistics + timings.  Never raises on a mismatch (one gpurun call should yield as much information as possible).

    python tools/gpu_diag.py [gemm] [attn] [elem] [tail]
"""
import math
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib, plan  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def stats(name, got, ref, tol=None):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    bad = torch.isnan(got).sum().item()
    msg = f"{name}: max_abs={err.max().item():.4e} mean_abs={err.mean().item():.4e} ref_max={denom:.3e} rel={err.max().item() / denom:.3e} nan={bad}"
    if tol is not None:
        msg += "  [OK]" if (err.max().item() / denom < tol and bad == 0) else "  [MISMATCH]"
    print(msg, flush=True)
    return err


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(name, fn):
    print(f"--- {name}", flush=True)
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:
        traceback.print_exc()
        print(f"--- {name} RAISED", flush=True)


def diag_gemm():
    def one(M, N, K, epi=lib.EPI_NONE, bias=False, res=False, out_dtype=torch.bfloat16, block_n=0, time_it=False):
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev) * 0.1 if bias else None
        r = (torch.randn(M, N, device=dev)).bfloat16() if res else None
        ref = a.float() @ w.float().t()
        if b is not None:
            ref = ref + b
        if epi == lib.EPI_GELU:
            ref = torch.nn.functional.gelu(ref)
        if epi == lib.EPI_SWIGLU:
            x = ref.view(M, N // 32, 2, 16)
            ref = (torch.nn.functional.silu(x[:, :, 0]) * x[:, :, 1]).reshape(M, N // 2)
        if r is not None:
            ref = ref + r.float()
        out = lib.gemm(a, w, bias=b, residual=r, epilogue=epi, out_dtype=out_dtype, block_n=block_n)
        torch.cuda.synchronize()
        err = stats(f"gemm M={M} N={N} K={K} epi={epi} bias={bias} res={res} out={out_dtype} bn={block_n}", out, ref, tol=2e-2)
        if err.max().item() / (ref.abs().max().item() + 1e-12) > 2e-2:
            e = err.view(M, -1)
            rows = (e.max(dim=1)[0] > 0.02 * ref.abs().max()).nonzero().flatten()
            cols = (e.max(dim=0)[0] > 0.02 * ref.abs().max()).nonzero().flatten()
            print(f"   bad rows: n={len(rows)} first={rows[:8].tolist()} last={rows[-4:].tolist()}; bad cols: n={len(cols)} first={cols[:8].tolist()} last={cols[-4:].tolist()}")
            print("   got[0,:8]", out[0, :8].float().tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())
        if time_it:
            ms = timeit(lambda: lib.gemm(a, w, out=out, bias=b, residual=r, epilogue=epi, block_n=block_n))
            print(f"   time {ms * 1e3:.1f} us  -> {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)

    run("gemm small exact tile", lambda: one(128, 256, 64))
    run("gemm K=128", lambda: one(128, 256, 128))
    run("gemm 2 tiles", lambda: one(256, 512, 256))
    run("gemm bn128", lambda: one(256, 512, 256, block_n=128))
    run("gemm bn32", lambda: one(256, 512, 256, block_n=32))
    run("gemm ragged M/K", lambda: one(300, 1280, 1176, bias=True))
    run("gemm small M auto", lambda: one(8, 2048, 2048, bias=True))
    run("gemm residual", lambda: one(1000, 1280, 1280, bias=True, res=True))
    run("gemm gelu", lambda: one(729, 5120, 5120, epi=lib.EPI_GELU, bias=True))
    run("gemm swiglu", lambda: one(1000, 6912, 1280, epi=lib.EPI_SWIGLU, bias=True))
    run("gemm swiglu bn32", lambda: one(8, 22016, 2048, epi=lib.EPI_SWIGLU))
    run("gemm f32 out ragged N", lambda: one(8, 151936, 2048, out_dtype=torch.float32))
    run("gemm f32 out bn256 ragged N", lambda: one(200, 151936 // 8, 2048, out_dtype=torch.float32, block_n=256))
    for M_ in (1, 8, 13, 16, 32):
        run(f"skinny M={M_} qkv", lambda M_=M_: one(M_, 2560, 2048, bias=True))
    run("skinny residual", lambda: one(8, 2048, 2048, res=True))
    run("skinny down", lambda: one(8, 2048, 11008, res=True, time_it=True))
    run("skinny swiglu", lambda: one(8, 22016, 2048, epi=lib.EPI_SWIGLU, time_it=True))
    run("skinny swiglu bias M=5", lambda: one(5, 6912, 1280, epi=lib.EPI_SWIGLU, bias=True))
    run("skinny lm head f32", lambda: one(8, 151936, 2048, out_dtype=torch.float32, time_it=True))
    run("skinny qkv perf", lambda: one(8, 2560, 2048, bias=True, time_it=True))
    run("skinny gelu M=20", lambda: one(20, 5120, 5120, epi=lib.EPI_GELU, bias=True))
    run("gemm perf vit qkv", lambda: one(23328, 3840, 1280, bias=True, time_it=True))
    run("gemm perf vit swiglu", lambda: one(23328, 6912, 1280, epi=lib.EPI_SWIGLU, bias=True, time_it=True))
    run("gemm perf vit down", lambda: one(23328, 1280, 3456, bias=True, res=True, time_it=True))
    run("gemm perf dec gate-up", lambda: one(7000, 22016, 2048, epi=lib.EPI_SWIGLU, time_it=True))
    run("gemm perf dec down", lambda: one(7000, 2048, 11008, res=True, time_it=True))
    run("gemm perf decode M=8 down", lambda: one(8, 2048, 11008, res=True, time_it=True))
    run("gemm perf decode M=8 gate-up", lambda: one(8, 22016, 2048, epi=lib.EPI_SWIGLU, time_it=True))


def ref_attention(q, k, v, row_lo, row_hi, scale):
    """q [Rq,H,D], k [Rk,Hkv,D], v [Rk,Hkv,Dv] fp32; mask from row_lo/hi."""
    Rq, H, D = q.shape
    Rk, Hkv, Dv = v.shape
    g = H // Hkv
    keys = torch.arange(Rk, device=q.device)[None, :]
    mask = (keys >= row_lo[:, None]) & (keys < row_hi[:, None])
    out = torch.empty(Rq, H, Dv, device=q.device)
    for h in range(H):
        s = (q[:, h].float() @ k[:, h // g].float().t()) * scale
        s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[:, h] = p @ v[:, h // g].float()
    return out


def diag_attn():
    def one(name, pl, n_rows, n_keys, H, Hkv, hdv, d_real, time_it=False):
        q = torch.zeros(n_rows, H, 128, device=dev, dtype=torch.bfloat16)
        k = torch.zeros(n_keys, Hkv, 128, device=dev, dtype=torch.bfloat16)
        q[:, :, :d_real] = torch.randn(n_rows, H, d_real, device=dev).bfloat16()
        k[:, :, :d_real] = torch.randn(n_keys, Hkv, d_real, device=dev).bfloat16()
        v = torch.randn(n_keys, Hkv, hdv, device=dev).bfloat16()
        ncol = (n_keys + 7) // 8 * 8
        vt = torch.zeros(Hkv * hdv, ncol, device=dev, dtype=torch.bfloat16)
        vt[:, :n_keys] = v.permute(1, 2, 0).reshape(Hkv * hdv, n_keys)
        out = torch.zeros(n_rows, H, hdv, device=dev, dtype=torch.bfloat16)
        tiles = torch.from_numpy(pl.tiles).to(dev)
        lo = torch.from_numpy(pl.row_lo).to(dev)
        hi = torch.from_numpy(pl.row_hi).to(dev)
        scale = d_real ** -0.5
        lib.attention(q, k, vt, out, tiles, lo, hi, scale, hdv)
        torch.cuda.synchronize()
        ref = ref_attention(q, k, v, lo, hi, scale)
        err = stats(f"attn {name} rows={n_rows} keys={n_keys} H={H}/{Hkv} hdv={hdv} tiles={len(pl.tiles)}", out, ref, tol=3e-2)
        if err.max().item() > 0.03 * ref.abs().max().item():
            e = err.amax(dim=(1, 2))
            rows = (e > 0.03 * ref.abs().max()).nonzero().flatten()
            print(f"   bad rows n={len(rows)} first={rows[:10].tolist()}")
            print("   got", out[0, 0, :6].float().tolist(), "\n   ref", ref[0, 0, :6].tolist())
        if time_it:
            ms = timeit(lambda: lib.attention(q, k, vt, out, tiles, lo, hi, scale, hdv))
            fl = 0
            for t in pl.tiles:
                fl += 2 * 128 * (math.ceil((t[3] - t[2]) / 128) * 128) * (128 + hdv) * H
            print(f"   time {ms * 1e3:.1f} us -> {fl / ms / 1e9:.1f} TFLOP/s (issued MMA flops)", flush=True)

    run("attn single tile full", lambda: one("1tile", plan.plan_from_segments(np.array([0, 128]), 128), 128, 128, 1, 1, 128, 128))
    run("attn single tile hdv80", lambda: one("1tile80", plan.plan_from_segments(np.array([0, 128]), 128), 128, 128, 2, 2, 80, 80))
    run("attn 2 kv blocks", lambda: one("2blk", plan.plan_from_segments(np.array([0, 256]), 256), 256, 256, 2, 2, 128, 128))
    run("attn 5 kv blocks ragged", lambda: one("5blk", plan.plan_from_segments(np.array([0, 600]), 600), 600, 600, 2, 1, 128, 128))
    vp = plan.vit_plan(np.array([[1, 54, 54]]))
    run("attn vit windows 756", lambda: one("win", vp.window_plan, vp.n, vp.n, 16, 16, 80, 80, time_it=True))
    run("attn vit full 756", lambda: one("full", vp.full_plan, vp.n, vp.n, 16, 16, 80, 80, time_it=True))
    vp2 = plan.vit_plan(np.array([[1, 16, 16], [1, 32, 64], [1, 74, 74]]))
    run("attn vit windows mixed", lambda: one("winmix", vp2.window_plan, vp2.n, vp2.n, 16, 16, 80, 80))
    run("attn vit full mixed", lambda: one("fullmix", vp2.full_plan, vp2.n, vp2.n, 16, 16, 80, 80))
    lens = [870, 333, 128, 1]
    Lmax = 1024
    qs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ks = np.arange(len(lens)) * Lmax

    def causal():
        pl = plan.plan_causal(lens, qs, ks)
        n_rows = int(sum(lens))
        n_keys = len(lens) * Lmax
        # keys only meaningful at cache rows; reference uses the full key buffer with masks
        one("causal", pl, n_rows, n_keys, 16, 2, 128, 128, time_it=True)

    run("attn causal gqa", causal)


def diag_elem():
    def rms():
        for D in (1280, 2048):
            x = torch.randn(1000, D, device=dev).bfloat16()
            w = 1 + 0.1 * torch.randn(D, device=dev)
            y = lib.rmsnorm(x, w, 1e-6)
            xf = x.float()
            ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
            stats(f"rmsnorm D={D}", y, ref, tol=1e-2)
    run("rmsnorm", rms)

    def gather():
        src = torch.randn(500, 1176, device=dev)
        idx = torch.randperm(500, device=dev).int()
        out = lib.gather_rows(src, idx)
        stats("gather f32->bf16", out, src[idx.long()].bfloat16(), tol=1e-6)
        s2 = torch.randn(300, 2048, device=dev).bfloat16()
        idx2 = torch.randint(0, 300, (77,), device=dev).int()
        stats("gather bf16", lib.gather_rows(s2, idx2), s2[idx2.long()], tol=1e-6)
    run("gather", gather)

    def embed():
        table = torch.randn(1000, 2048, device=dev).bfloat16()
        ids = torch.randint(0, 1000, (300,), device=dev)
        vis = torch.randn(64, 2048, device=dev).bfloat16()
        vidx = torch.full((300,), -1, device=dev, dtype=torch.int32)
        vidx[10:74] = torch.arange(64, device=dev, dtype=torch.int32)
        out = lib.embed_scatter(table, ids, vis, vidx)
        ref = table[ids].clone()
        ref[10:74] = vis
        stats("embed_scatter", out, ref, tol=1e-6)
    run("embed", embed)

    def vitrope():
        vp = plan.vit_plan(np.array([[1, 54, 54]]))
        n = vp.n
        inv = 1.0 / (10000.0 ** (torch.arange(0, 40, 2, dtype=torch.float) / 40))
        hp, wp = torch.from_numpy(vp.hpos).to(dev), torch.from_numpy(vp.wpos).to(dev)
        cos, sin = lib.vit_rope_table(hp, wp, inv.to(dev))
        fr = torch.cat([hp.float()[:, None] * inv.to(dev)[None], wp.float()[:, None] * inv.to(dev)[None]], 1)
        stats("vit rope cos", cos, fr.cos(), tol=1e-5)
        stats("vit rope sin", sin, fr.sin(), tol=1e-5)
        qkv = torch.randn(n, 3840, device=dev).bfloat16()
        qp = torch.zeros((n, 16, 128), device=dev, dtype=torch.bfloat16)
        kp = torch.zeros((n, 16, 128), device=dev, dtype=torch.bfloat16)
        npad = (n + 7) // 8 * 8
        vt = torch.zeros(16 * 80, npad, device=dev, dtype=torch.bfloat16)
        lib.vit_rope_split(qkv, cos, sin, qp, kp, vt)
        q, k, v = qkv.float().view(n, 3, 16, 80).unbind(1)
        c80 = torch.cat([fr.cos(), fr.cos()], 1)[:, None]
        s80 = torch.cat([fr.sin(), fr.sin()], 1)[:, None]
        rh = lambda x: torch.cat([-x[..., 40:], x[..., :40]], -1)
        stats("vit rope q", qp[:, :, :80], q * c80 + rh(q) * s80, tol=1e-2)
        stats("vit rope k", kp[:, :, :80], k * c80 + rh(k) * s80, tol=1e-2)
        print("   pad zero:", qp[:, :, 80:].abs().max().item(), kp[:, :, 80:].abs().max().item())
        stats("vit vt", vt[:, :n], v.permute(1, 2, 0).reshape(1280, n), tol=1e-6)
    run("vit rope", vitrope)

    def mrope():
        S = 777
        pos = torch.randint(0, 3000, (3, S), device=dev)
        inv = 1.0 / (1000000.0 ** (torch.arange(0, 128, 2, dtype=torch.float) / 128))
        cos, sin = lib.mrope_table(pos, inv.to(dev))
        fr = pos.float()[:, :, None] * inv.to(dev)[None, None]  # [3,S,64]
        sel = torch.cat([fr[0, :, :16], fr[1, :, 16:40], fr[2, :, 40:64]], -1)
        stats("mrope cos", cos, sel.cos(), tol=1e-4)
        stats("mrope sin", sin, sel.sin(), tol=1e-4)
        qkv = torch.randn(S, 2560, device=dev).bfloat16()
        orig = qkv.clone()
        rows = 1024
        cache_row = (torch.arange(S, device=dev) + 100).int()
        kc = torch.zeros(rows, 2, 128, device=dev, dtype=torch.bfloat16)
        vt = torch.zeros(256, rows, device=dev, dtype=torch.bfloat16)
        lib.mrope_kv(qkv, cos, sin, cache_row, kc, vt, 16, 2)
        c = torch.cat([sel.cos(), sel.cos()], -1)[:, None]
        s = torch.cat([sel.sin(), sel.sin()], -1)[:, None]
        rh = lambda x: torch.cat([-x[..., 64:], x[..., :64]], -1)
        q = orig[:, :2048].float().view(S, 16, 128)
        k = orig[:, 2048:2304].float().view(S, 2, 128)
        v = orig[:, 2304:].float().view(S, 2, 128)
        stats("mrope q", qkv[:, :2048].view(S, 16, 128), q * c + rh(q) * s, tol=1e-2)
        stats("mrope k cache", kc[100:100 + S], k * c + rh(k) * s, tol=1e-2)
        stats("mrope vt cache", vt[:, 100:100 + S], v.permute(1, 2, 0).reshape(256, S), tol=1e-6)
    run("mrope", mrope)


def diag_tail():
    def dec():
        B, Lmax = 4, 1024
        lens = torch.tensor([900, 17, 1, 513], device=dev, dtype=torch.int32)
        start = (torch.arange(B, device=dev) * Lmax).int()
        q = torch.randn(B, 2048, device=dev).bfloat16()
        kc = torch.randn(B * Lmax, 2, 128, device=dev).bfloat16()
        vt = torch.randn(256, B * Lmax, device=dev).bfloat16()
        out = torch.zeros(B, 2048, device=dev, dtype=torch.bfloat16)
        lib.decode_attention(q, kc, vt, start, lens, 1024, out, 16, 2, 128 ** -0.5)
        ref = torch.zeros(B, 16, 128, device=dev)
        for b in range(B):
            L = int(lens[b])
            kk = kc[b * Lmax:b * Lmax + L].float()
            vv = vt[:, b * Lmax:b * Lmax + L].float().view(2, 128, L)
            for h in range(16):
                s = (kk[:, h // 8] @ q[b, h * 128:(h + 1) * 128].float()) * 128 ** -0.5
                p = torch.softmax(s, 0)
                ref[b, h] = vv[h // 8] @ p
        stats("decode attention", out.view(B, 16, 128), ref, tol=2e-2)
    run("decode attn", dec)

    def am():
        lg = torch.randn(8, 151936, device=dev)
        stats("argmax f32", lib.argmax_rows(lg).float(), lg.argmax(-1).float(), tol=1e-9)
        lb = lg.bfloat16()
        got = lib.argmax_rows(lb)
        print("   argmax bf16 equal:", torch.equal(got, lb.float().argmax(-1)), got.tolist())
        labels = torch.randint(0, 151936, (8,), device=dev)
        lp, ent = lib.logprob_entropy(lg, labels)
        ls = torch.log_softmax(lg, -1)
        stats("logprob", lp, ls.gather(-1, labels[:, None]).squeeze(-1), tol=1e-5)
        stats("entropy", ent, -(ls.exp() * ls).sum(-1), tol=1e-5)
    run("argmax/logprob", am)

    def mask():
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(3000)
        lg = torch.randn(3, 3, 256, 256, generator=g)
        sc = torch.rand(3, 3, generator=g)
        out = lib.mask_epilogue(lg.to(dev), sc.to(dev))
        m = torch.zeros(756, 756, dtype=torch.bool)
        for o in range(3):
            up = F.interpolate(lg[o:o + 1], (756, 756), mode="bilinear", align_corners=False)[0]
            m |= (up > 0)[int(sc[o].argmax())]
        import cv2
        ref = cv2.resize(m.numpy().astype(np.uint8), (768, 768), interpolation=cv2.INTER_NEAREST)
        diff = (out.cpu().numpy() != ref).sum()
        print(f"mask epilogue: mismatching pixels {diff} / {768 * 768}  [{'OK' if diff == 0 else 'MISMATCH'}]")
        gt = (torch.rand(768, 768) < 0.3).to(torch.uint8) * 255
        iou, cnt = lib.mask_iou(out[None], gt[None].to(dev), 1.0, want_counts=True)
        p, gg = ref > 0, gt.numpy() > 0
        print("   iou", iou.item(), "ref", (p & gg).sum() / (p | gg).sum(), cnt.tolist(), (p & gg).sum(), (p | gg).sum())
        z = torch.zeros(2, 768, 768, dtype=torch.uint8, device=dev)
        print("   empty iou", lib.mask_iou(z, z, 1.0).tolist(), lib.mask_iou(z, z, 0.0).tolist())
        print("   empty-prompt mask sum", lib.mask_epilogue(torch.zeros(0, 3, 256, 256, device=dev), torch.zeros(0, 3, device=dev)).sum().item())
    run("mask", mask)


if __name__ == "__main__":
    lib.device_check()
    print(torch.cuda.get_device_name(0), "abi", lib.load().sr_abi_version(), flush=True)
    which = sys.argv[1:] or ["gemm", "attn", "elem", "tail"]
    t0 = time.time()
    if "gemm" in which:
        diag_gemm()
    if "attn" in which:
        diag_attn()
    if "elem" in which:
        diag_elem()
    if "tail" in which:
        diag_tail()
    print("done in", time.time() - t0, "s; launches", lib.launch_count)
