"""Bring-up / regression check of the persistent decode-step kernel (sr_decode_persistent.cu) against a plain fp32 torch
restatement of one decode step (HF:762-827 with use_cache, M-RoPE per mcore_adapter/.../modeling_qwen2_5_vl.py:21-41; bf16
rounding at the points where the kernel stores bf16) AND against the per-layer decode kernels, phase by phase (stop_after),
then a whole step, CUDA-graph replay and timing at the 3B geometry.  Synthetic packed weights (all paths read the same packed
tensors).  Usage: python tools/persistent_bringup.py [layers] [vocab]"""
import os
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib  # noqa: E402

dev = "cuda"
H, NQ, NKV, HD, I = 2048, 16, 2, 128, 11008
QKV = (NQ + 2 * NKV) * HD


def rnd_packed(n, k, std=0.02):
    return (torch.randn(n // 16, k // 32, 16, 32, device=dev) * std).bfloat16()


def make_layers(n_layers):
    from socioreasoner_b200.weights import _kslice_pack
    out = []
    for _ in range(n_layers):
        out.append(SimpleNamespace(
            qkv_wp=rnd_packed(QKV, H), qkv_b=(torch.randn(QKV, device=dev) * 0.02), ln1=1 + 0.1 * torch.randn(H, device=dev),
            o_wp=rnd_packed(H, H), ln2=1 + 0.1 * torch.randn(H, device=dev), gu_wp=rnd_packed(2 * I, H), down_wp=rnd_packed(H, I)))
        L = out[-1]   # the persistent kernel reads the down projection K-sliced (same matrix, third layout)
        L.down_wk = _kslice_pack(L.down_wp.permute(0, 2, 1, 3).reshape(H, I).contiguous())
    return out


def err(name, got, ref):
    got, ref = got.float(), ref.float()
    d = (got - ref).abs().max().item()
    r = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-20)).item()
    ok = torch.isfinite(got).all().item() and r < 2e-2
    print(f"  {name:34s} max|d|={d:.4e} rel_rms={r:.3e} {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


class State:
    def __init__(self, B, n_layers, l_max, lens, vocab):
        g = torch.Generator(device=dev).manual_seed(1)
        self.B, self.l_max = B, l_max
        self.x0 = torch.randn(B, H, device=dev, generator=g).bfloat16()
        self.kc0 = (torch.randn(n_layers, B * l_max, NKV, HD, device=dev, generator=g)).bfloat16()
        self.vt0 = (torch.randn(n_layers, NKV * HD, B * l_max, device=dev, generator=g)).bfloat16()
        self.kv_start = (torch.arange(B, device=dev) * l_max).int()
        self.kv_len = torch.tensor(lens, device=dev, dtype=torch.int32)            # includes the new token
        self.cache_row = (self.kv_start + self.kv_len - 1).int()
        pos = (self.kv_len.long() + 3)[None].repeat(3, 1).contiguous()
        inv = (1.0 / (1e6 ** (torch.arange(0, HD, 2, dtype=torch.float32) / HD))).to(dev)
        self.cos, self.sin = lib.mrope_table(pos, inv, (16, 24, 24))
        self.final_norm = 1 + 0.1 * torch.randn(H, device=dev, generator=g)
        self.lm_head_p = rnd_packed(vocab, H)
        self.vocab = vocab


def unpack(wp, n, k):
    """[N/16][K/32][16][32] streaming layout -> [N, K] fp32."""
    return wp.permute(0, 2, 1, 3).reshape(n, k).float()


def _bf(t):
    return t.to(torch.bfloat16).float()


def reference_step_torch(S, layers):
    """fp32 torch restatement of the step (the parity reference of the GPU tests)."""
    B = S.B
    x = S.x0.float()
    kc, vt = S.kc0.clone(), S.vt0.clone()
    cos, sin = S.cos.float(), S.sin.float()            # [B, 64]
    inter = {}

    def norm(v, w):
        return _bf(v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-6) * w)

    def rope(t):                                       # [B, heads, 128]
        x1, x2 = t[..., :64], t[..., 64:]
        c, s_ = cos[:, None, :], sin[:, None, :]
        return _bf(torch.cat([x1 * c - x2 * s_, x2 * c + x1 * s_], -1))

    for li, L in enumerate(layers):
        qkv = _bf(norm(x, L.ln1) @ unpack(L.qkv_wp, QKV, H).T + L.qkv_b)
        if li == 0:
            inter["qkv"] = qkv.clone()
        q = rope(qkv[:, :NQ * HD].view(B, NQ, HD))
        k = rope(qkv[:, NQ * HD:(NQ + NKV) * HD].view(B, NKV, HD))
        v = qkv[:, (NQ + NKV) * HD:].view(B, NKV, HD)
        ao = torch.zeros(B, NQ, HD, device=dev)
        for b in range(B):
            row, st, ln = int(S.cache_row[b]), int(S.kv_start[b]), int(S.kv_len[b])
            kc[li][row] = k[b].to(torch.bfloat16)
            vt[li][:, row] = v[b].reshape(-1).to(torch.bfloat16)
            for kvh in range(NKV):
                keys = kc[li][st:st + ln, kvh].float()                            # [len, 128]
                vals = vt[li][kvh * HD:(kvh + 1) * HD, st:st + ln].float()       # [128, len]
                qh = q[b, kvh * (NQ // NKV):(kvh + 1) * (NQ // NKV)]              # [8, 128]
                pr = torch.softmax(qh @ keys.T * HD ** -0.5, -1)
                ao[b, kvh * (NQ // NKV):(kvh + 1) * (NQ // NKV)] = pr @ vals.T
        ao = _bf(ao.view(B, H))
        if li == 0:
            inter["ao"] = ao.clone()
        x = _bf(x + ao @ unpack(L.o_wp, H, H).T)
        if li == 0:
            inter["x_o"] = x.clone()
        gu = (norm(x, L.ln2) @ unpack(L.gu_wp, 2 * I, H).T).view(B, I // 16, 2, 16)
        act = _bf((torch.nn.functional.silu(gu[:, :, 0]) * gu[:, :, 1]).reshape(B, I))
        if li == 0:
            inter["act"] = act.clone()
        x = _bf(x + act @ unpack(L.down_wp, H, I).T)
        if li == 0:
            inter["x_down"] = x.clone()
    logits = norm(x, S.final_norm) @ unpack(S.lm_head_p, S.vocab, H).T
    inter.update(x=x, logits=logits, kc=kc, vt=vt)
    return inter


def reference_step(S, layers, upto_layers=None):
    """Per-layer kernels (the path generate_greedy used before): returns intermediates of layer 0 and the final logits."""
    x = S.x0.clone()
    kc, vt = S.kc0.clone(), S.vt0.clone()
    B = S.B
    qkv = torch.empty(B, QKV, device=dev, dtype=torch.bfloat16)
    ao = torch.empty(B, H, device=dev, dtype=torch.bfloat16)
    act = torch.empty(B, I, device=dev, dtype=torch.bfloat16)
    inter = {}
    for li, L in enumerate(layers):
        lib.gemm_skinny_packed(x, L.qkv_wp, QKV, H, out=qkv, bias=L.qkv_b, norm_weight=L.ln1, norm_eps=1e-6)
        if li == 0:
            inter["qkv"] = qkv.clone()
        lib.mrope_kv(qkv, S.cos, S.sin, S.cache_row, kc[li], vt[li], NQ, NKV)
        lib.decode_attention(qkv, kc[li], vt[li], S.kv_start, S.kv_len, S.l_max, ao, NQ, NKV, HD ** -0.5)
        if li == 0:
            inter["ao"] = ao.clone()
        lib.gemm_skinny_packed(ao, L.o_wp, H, H, out=x, residual=x)
        if li == 0:
            inter["x_o"] = x.clone()
        lib.gemm_skinny_packed(x, L.gu_wp, 2 * I, H, out=act, epilogue=lib.EPI_SWIGLU, norm_weight=L.ln2, norm_eps=1e-6)
        if li == 0:
            inter["act"] = act.clone()
        lib.gemm_skinny_packed(act, L.down_wp, H, I, out=x, residual=x)
        if li == 0:
            inter["x_down"] = x.clone()
    y = lib.rmsnorm(x, S.final_norm, 1e-6)
    logits = lib.gemm_skinny_packed(y, S.lm_head_p, S.vocab, H, out_dtype=torch.float32)
    inter.update(x=x, logits=logits, kc=kc, vt=vt)
    return inter


class Persistent:
    def __init__(self, S, layers):
        self.S, self.layers = S, layers
        self.table = lib.decode_layer_table(layers, dev)
        B = S.B
        self.qkv = torch.zeros(B, QKV, device=dev, dtype=torch.bfloat16)
        self.ao = torch.zeros(B, H, device=dev, dtype=torch.bfloat16)
        self.act = torch.zeros(B, I, device=dev, dtype=torch.bfloat16)
        self.logits = torch.zeros(B, S.vocab, device=dev, dtype=torch.float32)
        self.ws = lib.decode_persistent_workspace(NQ, NKV, I, dev)
        self.x = S.x0.clone()
        self.kc, self.vt = S.kc0.clone(), S.vt0.clone()

    def reset(self):
        self.x.copy_(self.S.x0); self.kc.copy_(self.S.kc0); self.vt.copy_(self.S.vt0)
        for t in (self.qkv, self.ao, self.act, self.logits):
            t.zero_()

    def run(self, stop_after=0):
        S = self.S
        lib.decode_step_persistent(self.table, S.lm_head_p, S.final_norm, S.vocab, self.x, self.qkv, self.ao, self.act, self.logits,
                                   S.cos, S.sin, S.cache_row, S.kv_start, S.kv_len, self.kc, self.vt, NQ, NKV, I, 1e-6, HD ** -0.5,
                                   self.ws, stop_after)


def check_case(B, n_layers, lens, vocab, l_max=1024, phases=True):
    print(f"== case B={B} layers={n_layers} lens={lens} vocab={vocab}", flush=True)
    torch.manual_seed(0)
    layers = make_layers(n_layers)
    S = State(B, n_layers, l_max, lens, vocab)
    lib.decode_persistent_tuning(flags=3)  # dump plain intermediates (x / qkv / ao / act) next to the LL buffers
    ref = reference_step_torch(S, layers)
    ref_k = reference_step(S, layers)
    torch.cuda.synchronize()
    ok = err("per-layer kernels vs torch fp32: logits", ref_k["logits"], ref["logits"])
    P = Persistent(S, layers)
    if phases:
        for k, (name, got_fn, refkey) in enumerate([
                ("phase1 qkv", lambda: P.qkv, "qkv"), ("phase2 attention out", lambda: P.ao, "ao"),
                ("phase3 x after o-proj", lambda: P.x, "x_o"), ("phase4 act (SwiGLU)", lambda: P.act, "act"),
                ("phase5 x after down", lambda: P.x, "x_down")], start=1):
            P.reset()
            P.run(stop_after=k)
            torch.cuda.synchronize()
            ok &= err(name, got_fn(), ref[refkey])
            if k == 2:
                ok &= err("  K cache (layer 0)", P.kc[0], ref["kc"][0])
                ok &= err("  V^T cache (layer 0)", P.vt[0], ref["vt"][0])
    P.reset()
    P.run(0)
    torch.cuda.synchronize()
    ok &= err("full step: hidden", P.x, ref["x"])
    ok &= err("full step: logits", P.logits, ref["logits"])
    ok &= err("full step: K cache (all)", P.kc, ref["kc"])
    ok &= err("full step: V^T cache (all)", P.vt, ref["vt"])
    am = (P.logits.argmax(-1) == ref["logits"].argmax(-1)).float().mean().item()
    print(f"  argmax agreement {am:.2f}", flush=True)
    # second run on the same object (barrier counter / tickets re-armed)
    P.reset(); P.run(0); torch.cuda.synchronize()
    ok &= err("full step again: logits", P.logits, ref["logits"])
    lib.decode_persistent_tuning(flags=1)
    P.reset(); P.run(0); torch.cuda.synchronize()
    ok &= err("full step without dumps: logits", P.logits, ref["logits"])
    return ok, S, layers, P, ref


def graph_check(P, ref):
    print("== CUDA graph capture + replay", flush=True)
    P.reset(); P.run(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    P.reset()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        P.run(0)
    P.reset()
    g.replay(); torch.cuda.synchronize()
    ok = err("graph replay: logits", P.logits, ref["logits"])
    P.reset(); g.replay(); torch.cuda.synchronize()
    ok &= err("graph replay 2: logits", P.logits, ref["logits"])
    return ok, g


def timing(n_layers=36, vocab=151936, B=8, lens=None, iters=20):
    print(f"== timing layers={n_layers} vocab={vocab} B={B}", flush=True)
    torch.manual_seed(0)
    lens = lens or [900 - 7 * i for i in range(B)]
    layers = make_layers(n_layers)
    S = State(B, n_layers, 1024, lens, vocab)
    P = Persistent(S, layers)
    wbytes = n_layers * (QKV * H + H * H + 2 * I * H + H * I) * 2 + vocab * H * 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        P.run(0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        P.run(0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"  persistent (stream launches): {ms:.3f} ms/step  {wbytes / ms / 1e9:.2f} TB/s of weights ({wbytes / 1e9:.2f} GB)", flush=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        P.run(0)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"  persistent (graph replay):    {ms:.3f} ms/step  {wbytes / ms / 1e9:.2f} TB/s of weights", flush=True)
    # per-phase cumulative timing: stop_after = k phases of layer 0..1, and the no-LM-head prefix
    for k in (1, 2, 3, 4, 5, 10, 5 * n_layers):
        for _ in range(2):
            P.run(k)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            P.run(k)
        e1.record(); torch.cuda.synchronize()
        print(f"    first {k:4d} phases: {e0.elapsed_time(e1) / iters * 1e3:9.1f} us", flush=True)
    # the per-layer kernels on the same data, captured in a graph (what generate_greedy replayed before)
    x = S.x0.clone(); kc, vt = S.kc0.clone(), S.vt0.clone()
    qkv = torch.empty(B, QKV, device=dev, dtype=torch.bfloat16); ao = torch.empty(B, H, device=dev, dtype=torch.bfloat16)
    act = torch.empty(B, I, device=dev, dtype=torch.bfloat16); y = torch.empty(B, H, device=dev, dtype=torch.bfloat16)
    logits = torch.empty(B, vocab, device=dev, dtype=torch.float32)

    def old():
        for li, L in enumerate(layers):
            lib.gemm_skinny_packed(x, L.qkv_wp, QKV, H, out=qkv, bias=L.qkv_b, norm_weight=L.ln1, norm_eps=1e-6)
            lib.mrope_kv(qkv, S.cos, S.sin, S.cache_row, kc[li], vt[li], NQ, NKV)
            lib.decode_attention(qkv, kc[li], vt[li], S.kv_start, S.kv_len, S.l_max, ao, NQ, NKV, HD ** -0.5)
            lib.gemm_skinny_packed(ao, L.o_wp, H, H, out=x, residual=x)
            lib.gemm_skinny_packed(x, L.gu_wp, 2 * I, H, out=act, epilogue=lib.EPI_SWIGLU, norm_weight=L.ln2, norm_eps=1e-6)
            lib.gemm_skinny_packed(act, L.down_wp, H, I, out=x, residual=x)
        lib.rmsnorm(x, S.final_norm, 1e-6, out=y)
        lib.gemm_skinny_packed(y, S.lm_head_p, vocab, H, out=logits, out_dtype=torch.float32)
    old(); torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        old()
    for _ in range(2):
        g2.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        g2.replay()
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"  per-layer kernels (graph replay): {ms2:.3f} ms/step  {wbytes / ms2 / 1e9:.2f} TB/s of weights", flush=True)


if __name__ == "__main__":
    n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 36
    vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 151936
    lib.device_check()
    t0 = time.time()
    all_ok = True
    ok, S, layers, P, ref = check_case(8, 2, [900, 897, 640, 512, 129, 128, 9, 1], 8192)
    all_ok &= ok
    ok, _g = graph_check(P, ref)
    all_ok &= ok
    del S, layers, P, ref, _g
    ok, *_ = check_case(3, 3, [300, 17, 1000], 4096, phases=False)
    all_ok &= ok
    del _
    # contexts beyond 128 keys per split (multi-pass attention loop)
    ok, *_ = check_case(8, 2, [1500, 1400, 1300, 1200, 1153, 1152, 300, 1], 8192, l_max=1536, phases=False)
    all_ok &= ok
    print(f"long-context (multi-pass) case: {'OK' if ok else 'MISMATCH'}", flush=True)
    torch.cuda.empty_cache()
    print(f"correctness: {'ALL OK' if all_ok else 'FAILURES'} ({time.time() - t0:.1f}s)", flush=True)
    timing(n_layers, vocab)
