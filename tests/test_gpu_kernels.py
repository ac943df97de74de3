"""GPU (-m gpu): every kernel, called through the C ABI (socioreasoner_b200.lib -> libsocioreasoner_b200.so), against
fp32 CPU restatements of the reference op (oracle/host_ref.py, torch fp32 for the floating-point kernels).
Tolerances: bf16 storage => 2^-8 relative per rounding; stated per test.  Integer / index results are bit-exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hf_oracle as O
from oracle import host_ref
from socioreasoner_b200 import lib, plan

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_max(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g) * scale


# ---------------------------------------------------------------------------------------------------- GEMM
def gemm_ref(a, w, bias, res, epi):
    ref = a.float() @ w.float().t()
    if bias is not None:
        ref = ref + bias
    if epi == lib.EPI_GELU:
        ref = F.gelu(ref)  # exact erf GELU (HF merger, HF:133-146)
    if epi == lib.EPI_SWIGLU:
        x = ref.view(ref.shape[0], -1, 2, 16)
        ref = (F.silu(x[:, :, 0]) * x[:, :, 1]).reshape(ref.shape[0], -1)
    if res is not None:
        ref = ref + res.float()
    return ref


GEMM_CASES = [
    # M, N, K, epi, bias, res, out_f32, block_n
    (128, 256, 64, 0, False, False, False, 0),           # one exact tile
    (300, 1280, 1176, 0, True, False, False, 0),         # patch embed: ragged M, K tail (1176 = 18*64 + 24)
    (1000, 3840, 1280, 0, True, False, False, 0),        # ViT qkv
    (1000, 1280, 1280, 0, True, True, False, 0),         # ViT proj + residual
    (777, 6912, 1280, 2, True, False, False, 0),         # ViT SwiGLU (I padded 3420 -> 3456)
    (777, 1280, 3456, 0, True, True, False, 0),          # ViT down
    (365, 5120, 5120, 3, True, False, False, 0),         # merger fc1 + GELU
    (365, 2048, 5120, 0, True, False, False, 0),         # merger fc2
    (900, 2560, 2048, 0, True, False, False, 0),         # decoder qkv
    (900, 22016, 2048, 2, False, False, False, 0),       # decoder gate/up
    (900, 2048, 11008, 0, False, True, False, 0),        # decoder down
    (130, 151936 // 16, 2048, 0, False, False, True, 0),  # lm head fp32, N tail (9496 = 37*256 + 24)
    (256, 512, 256, 0, False, False, False, 128),
    (256, 512, 256, 0, True, True, False, 32),
    (1, 2560, 2048, 0, True, False, False, 0),           # skinny (decode) path
    (8, 2048, 2048, 0, False, True, False, 0),
    (8, 22016, 2048, 2, False, False, False, 0),
    (8, 2048, 11008, 0, False, True, False, 0),
    (13, 6912, 1280, 2, True, False, False, 0),
    (32, 5120, 5120, 3, True, False, False, 0),
    (8, 151936, 2048, 0, False, False, True, 0),         # full-vocab LM head, fp32 logits
    (40, 2048, 2048, 0, True, False, False, 0),          # M in (32, 64]: tcgen05 BLOCK_N=32 tiles
]


@pytest.mark.parametrize("M,N,K,epi,bias,res,f32,bn", GEMM_CASES)
def test_gemm(M, N, K, epi, bias, res, f32, bn):
    a = rnd(M, K, scale=0.5).bfloat16()
    w = rnd(N, K, scale=0.05, seed=1).bfloat16()
    b = rnd(N, scale=0.1, seed=2) if bias else None
    n_out = N // 2 if epi == lib.EPI_SWIGLU else N
    r = rnd(M, n_out, seed=3).bfloat16() if res else None
    out = lib.gemm(a.to(DEV), w.to(DEV), bias=b.to(DEV) if bias else None, residual=r.to(DEV) if res else None,
                   epilogue=epi, out_dtype=torch.float32 if f32 else torch.bfloat16, block_n=bn)
    ref = gemm_ref(a, w, b, r, epi)
    # fp32 accumulate; bf16 output rounding 2^-9 relative to the row's largest value is the only error source
    assert rel_max(out, ref) < (2e-5 if f32 else 6e-3)


@pytest.mark.parametrize("M,N,K,epi,f32", [(8, 2560, 2048, 0, False), (8, 22016, 2048, 2, False), (3, 151936, 2048, 0, True),
                                             (16, 2048, 2048, 0, False), (29, 2560, 2048, 0, False)])
def test_skinny_gemm_with_fused_rmsnorm(M, N, K, epi, f32):
    """Decode step: RMSNorm (HF:57-71) folded into the weight-streaming GEMM == rmsnorm followed by the linear layer."""
    a = rnd(M, K, scale=1.5).bfloat16()
    w = rnd(N, K, scale=0.05, seed=1).bfloat16()
    nw = 1 + 0.1 * rnd(K, seed=2)
    b = rnd(N, scale=0.1, seed=3) if epi == 0 and not f32 else None
    out = lib.gemm(a.to(DEV), w.to(DEV), bias=b.to(DEV) if b is not None else None, epilogue=epi,
                   out_dtype=torch.float32 if f32 else torch.bfloat16, norm_weight=nw.to(DEV), norm_eps=1e-6)
    af = a.float()
    normed = nw * (af * torch.rsqrt(af.pow(2).mean(-1, keepdim=True) + 1e-6))
    ref = gemm_ref(normed, w, b, None, epi)
    assert rel_max(out, ref) < 8e-3   # one extra bf16 rounding of (x * w_norm) before the MMA


@pytest.mark.parametrize("M,N,K,epi,norm", [(8, 2048, 11008, 0, False), (8, 22016, 2048, 2, True), (5, 2560, 2048, 0, True)])
def test_skinny_gemm_packed_layout_is_bit_identical(M, N, K, epi, norm):
    """The streaming weight layout is a pure re-layout: same products in the same order => identical bits (LDG kernel);
    the bulk-copy variant splits K over 8 instead of 16 warps => equal up to fp32 re-association + one bf16 rounding."""
    a = rnd(M, K, scale=0.7).bfloat16().to(DEV)
    w = rnd(N, K, scale=0.05, seed=1).bfloat16().to(DEV)
    nw = (1 + 0.1 * rnd(K, seed=2)).to(DEV) if norm else None
    ref = lib.gemm(a, w, epilogue=epi, norm_weight=nw)
    wp = lib.pack_skinny(w)
    lib.load().sr_set_gemv_bulk(0)
    got = lib.gemm_skinny_packed(a, wp, N, K, epilogue=epi, norm_weight=nw)
    assert torch.equal(ref, got)
    lib.load().sr_set_gemv_bulk(1)
    try:
        bulk = lib.gemm_skinny_packed(a, wp, N, K, epilogue=epi, norm_weight=nw)
    finally:
        lib.load().sr_set_gemv_bulk(0)
    assert rel_max(bulk, ref) < 8e-3   # one bf16 ulp of the largest element = 3.9e-3


@pytest.mark.parametrize("M,N,K,epi,f32,bias,res", [(8, 2560, 2048, 0, False, True, False), (1, 2048, 2048, 0, False, False, True),
                                                  (8, 2048, 11008, 0, False, False, True), (7, 22016, 2048, 2, False, False, False),
                                                  (8, 151936, 2048, 0, True, False, False), (3, 5120, 5120, 3, False, True, False),
                                                  (8, 2048, 2080, 0, False, False, False)])
def test_skinny_gemm_bulk_engine(M, N, K, epi, f32, bias, res):
    """cp.async.bulk streaming variant vs the fp32 reference (ragged last stage: K = 2080 = 65 chunks)."""
    a = rnd(M, K, scale=0.5).bfloat16()
    w = rnd(N, K, scale=0.05, seed=1).bfloat16()
    b = rnd(N, scale=0.1, seed=2) if bias else None
    n_out = N // 2 if epi == lib.EPI_SWIGLU else N
    r = rnd(M, n_out, seed=3).bfloat16() if res else None
    lib.load().sr_set_gemv_bulk(1)
    try:
        out = lib.gemm_skinny_packed(a.to(DEV), lib.pack_skinny(w.to(DEV)), N, K, bias=b.to(DEV) if bias else None,
                                     residual=r.to(DEV) if res else None, epilogue=epi, out_dtype=torch.float32 if f32 else torch.bfloat16)
    finally:
        lib.load().sr_set_gemv_bulk(0)
    assert rel_max(out, gemm_ref(a, w, b, r, epi)) < (2e-5 if f32 else 6e-3)


@pytest.mark.parametrize("M,N,K,res", [(8, 2048, 11008, True), (3, 2048, 2048, False), (8, 2752, 8192, True)])
def test_skinny_gemm_split_k(M, N, K, res):
    a = rnd(M, K, scale=0.5).bfloat16()
    w = rnd(N, K, scale=0.05, seed=1).bfloat16()
    r = rnd(M, N, seed=3).bfloat16() if res else None
    wp = lib.pack_skinny(w.to(DEV))
    for _ in range(3):   # repeated launches: the in-kernel ticket re-arm must hold
        out = lib.gemm_skinny_packed(a.to(DEV), wp, N, K, residual=r.to(DEV) if res else None, split_k=True)
        assert rel_max(out, gemm_ref(a, w, None, r, 0)) < 6e-3


def test_gemm_linearity_and_row_independence():
    """Size-independent properties at a production shape: rows are independent (batch invariance is bit-exact because
    every output element sees the same k-block order) and f(2a) == 2 f(a) exactly for a power-of-two scale."""
    M, N, K = 23328, 1280, 1280
    a = rnd(M, K, scale=0.5).bfloat16().to(DEV)
    w = rnd(N, K, scale=0.05, seed=1).bfloat16().to(DEV)
    full = lib.gemm(a, w)
    part = lib.gemm(a[5000:5300].contiguous(), w)
    assert torch.equal(full[5000:5300], part)
    assert torch.equal(lib.gemm((a * 2).contiguous(), w), full * 2)


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(8, 100, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(16, 100, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError):  # K row stride not a multiple of 8 elements (TMA 16-byte rule)
        lib.gemm(a, w, block_n=32)
    with pytest.raises(RuntimeError):
        lib.gemm(torch.zeros(8, 64, dtype=torch.float32, device=DEV), torch.zeros(16, 64, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(RuntimeError):
        lib.gemm(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(16, 64, dtype=torch.bfloat16))  # CPU tensors: no fallback


# ---------------------------------------------------------------------------------------------------- attention
def attention_ref(q, k, v, lo, hi, scale):
    Rq, H, _ = q.shape
    Rk, Hkv, Dv = v.shape
    g = H // Hkv
    keys = torch.arange(Rk)[None, :]
    mask = (keys >= lo[:, None]) & (keys < hi[:, None])
    out = torch.empty(Rq, H, Dv)
    for h in range(H):
        s = (q[:, h].float() @ k[:, h // g].float().t()) * scale
        p = torch.softmax(s.masked_fill(~mask, float("-inf")), dim=-1)
        out[:, h] = p @ v[:, h // g].float()
    return out


def run_attention(pl, n_rows, n_keys, H, Hkv, hdv, d_real, k_gain=None):
    q = torch.zeros(n_rows, H, 128, dtype=torch.bfloat16)
    k = torch.zeros(n_keys, Hkv, 128, dtype=torch.bfloat16)
    q[:, :, :d_real] = rnd(n_rows, H, d_real).bfloat16()
    kk = rnd(n_keys, Hkv, d_real, seed=1)
    if k_gain is not None:   # per-key gain on the key vectors: scores whose spread grows along the key axis
        kk = kk * k_gain.view(n_keys, 1, 1)
    k[:, :, :d_real] = kk.bfloat16()
    v = rnd(n_keys, Hkv, hdv, seed=2).bfloat16()
    ncol = (n_keys + 7) // 8 * 8
    vt = torch.zeros(Hkv * hdv, ncol, dtype=torch.bfloat16)
    vt[:, :n_keys] = v.permute(1, 2, 0).reshape(Hkv * hdv, n_keys)
    out = torch.zeros(n_rows, H, hdv, dtype=torch.bfloat16, device=DEV)
    lo, hi = torch.from_numpy(pl.row_lo), torch.from_numpy(pl.row_hi)
    lib.attention(q.to(DEV), k.to(DEV), vt.to(DEV), out, torch.from_numpy(pl.tiles).to(DEV), lo.to(DEV), hi.to(DEV), d_real ** -0.5, hdv)
    ref = attention_ref(q, k, v, lo.long(), hi.long(), d_real ** -0.5)
    return rel_max(out, ref)


@pytest.mark.parametrize("grid", [[[1, 54, 54]], [[1, 54, 54]] * 2, [[1, 16, 16], [1, 32, 64], [1, 30, 38]], [[1, 2, 2]]])
def test_attention_vit_windows_and_full(grid):
    vp = plan.vit_plan(np.array(grid))
    # P is rounded to bf16 before the PV MMA (as flash-attn does): 2^-8 relative on the largest output
    assert run_attention(vp.window_plan, vp.n, vp.n, 4, 4, 80, 80) < 8e-3
    assert run_attention(vp.full_plan, vp.n, vp.n, 2, 2, 80, 80) < 8e-3


def test_attention_causal_gqa_ragged():
    lens, L = [870, 333, 128, 1, 129], 1024
    qs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ks = np.arange(len(lens)) * L
    assert run_attention(plan.plan_causal(lens, qs, ks), int(sum(lens)), len(lens) * L, 16, 2, 128, 128) < 8e-3


@pytest.mark.parametrize("hdv,H,Hkv", [(80, 2, 2), (128, 4, 2)])
def test_attention_moving_maximum(hdv, H, Hkv):
    """The product attention kernel keeps a LAZY softmax reference point (it moves only when a block's maximum exceeds it by 2^8)
    and rescales O in tensor memory when it does.  Random scores never trigger that, so: keys whose magnitude grows 40-fold along
    the sequence (every later block raises the row maximum by far more than 2^8), bidirectional over 5 blocks and causal."""
    n = 640
    gain = torch.linspace(0.25, 10.0, n)
    full = plan.plan_from_segments(np.array([0, n]), n)
    assert run_attention(full, n, n, H, Hkv, hdv, hdv, k_gain=gain) < 8e-3
    cp = plan.plan_causal([n], [0], [0])
    assert run_attention(cp, n, n, H, Hkv, hdv, hdv, k_gain=gain) < 8e-3
    # and the other way round (the maximum is set by the first block and never moves again)
    assert run_attention(full, n, n, H, Hkv, hdv, hdv, k_gain=gain.flip(0)) < 8e-3


def test_decode_attention():
    B, L = 4, 1024
    lens = torch.tensor([900, 17, 1, 513], dtype=torch.int32)
    start = (torch.arange(B) * L).int()
    q = rnd(B, 2048).bfloat16()
    kc = rnd(B * L, 2, 128, seed=1).bfloat16()
    vt = rnd(256, B * L, seed=2).bfloat16()
    out = torch.zeros(B, 2048, dtype=torch.bfloat16, device=DEV)
    lib.decode_attention(q.to(DEV), kc.to(DEV), vt.to(DEV), start.to(DEV), lens.to(DEV), L, out, 16, 2, 128 ** -0.5)
    ref = torch.zeros(B, 16, 128)
    for b in range(B):
        n = int(lens[b])
        kk = kc[b * L:b * L + n].float()
        vv = vt[:, b * L:b * L + n].float().view(2, 128, n)
        for h in range(16):
            p = torch.softmax((kk[:, h // 8] @ q[b, h * 128:(h + 1) * 128].float()) * 128 ** -0.5, 0)
            ref[b, h] = vv[h // 8] @ p
    assert rel_max(out.view(B, 16, 128), ref) < 6e-3


# ---------------------------------------------------------------------------------------------------- element-wise
@pytest.mark.parametrize("D", [1280, 2048])
def test_rmsnorm(D):
    x = rnd(1003, D).bfloat16()
    w = 1 + 0.1 * rnd(D, seed=1)
    y = lib.rmsnorm(x.to(DEV), w.to(DEV), 1e-6)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))  # HF:57-71 in fp32
    assert rel_max(y, ref) < 4e-3


def test_gather_and_embed_scatter_bit_exact():
    src = rnd(500, 1176)
    idx = torch.randperm(500, generator=torch.Generator().manual_seed(0)).int()
    assert torch.equal(lib.gather_rows(src.to(DEV), idx.to(DEV)).cpu(), src[idx.long()].bfloat16())
    table = rnd(1000, 2048).bfloat16()
    ids = torch.randint(0, 1000, (300,), generator=torch.Generator().manual_seed(1))
    vis = rnd(64, 2048, seed=5).bfloat16()
    vidx = torch.full((300,), -1, dtype=torch.int32)
    vidx[10:74] = torch.randperm(64, generator=torch.Generator().manual_seed(2)).int()
    out = lib.embed_scatter(table.to(DEV), ids.to(DEV), vis.to(DEV), vidx.to(DEV)).cpu()
    ref = table[ids].clone()
    ref[10:74] = vis[vidx[10:74].long()]
    assert torch.equal(out, ref)


def test_vit_rope_tables_and_split():
    vp = plan.vit_plan(np.array([[1, 54, 54], [1, 16, 20]]))
    n = vp.n
    inv = 1.0 / (10000.0 ** (torch.arange(0, 40, 2, dtype=torch.float) / 40))
    hp, wp = torch.from_numpy(vp.hpos), torch.from_numpy(vp.wpos)
    cos, sin = lib.vit_rope_table(hp.to(DEV), wp.to(DEV), inv.to(DEV))
    fr = torch.cat([hp.float()[:, None] * inv[None], wp.float()[:, None] * inv[None]], 1)
    assert (cos.cpu() - fr.cos()).abs().max() < 2e-6 and (sin.cpu() - fr.sin()).abs().max() < 2e-6
    qkv = rnd(n, 3840).bfloat16()
    qp = torch.zeros((n, 16, 128), dtype=torch.bfloat16, device=DEV)   # caller zero-fills the 80..127 padding once
    kp = torch.zeros((n, 16, 128), dtype=torch.bfloat16, device=DEV)
    vt = torch.zeros(1280, (n + 7) // 8 * 8, dtype=torch.bfloat16, device=DEV)
    lib.vit_rope_split(qkv.to(DEV), cos, sin, qp, kp, vt)
    q, k, v = qkv.float().view(n, 3, 16, 80).unbind(1)
    c80, s80 = torch.cat([fr.cos()] * 2, 1)[:, None], torch.cat([fr.sin()] * 2, 1)[:, None]
    rh = lambda x: torch.cat([-x[..., 40:], x[..., :40]], -1)  # HF:149-153
    assert rel_max(qp[:, :, :80], q * c80 + rh(q) * s80) < 4e-3
    assert rel_max(kp[:, :, :80], k * c80 + rh(k) * s80) < 4e-3
    assert qp[:, :, 80:].abs().max() == 0 and kp[:, :, 80:].abs().max() == 0
    assert torch.equal(vt[:, :n].cpu(), v.permute(1, 2, 0).reshape(1280, n).bfloat16())


def test_mrope_table_apply_and_cache():
    S = 777
    pos = torch.randint(0, 5000, (3, S), generator=torch.Generator().manual_seed(0))
    inv = 1.0 / (1000000.0 ** (torch.arange(0, 128, 2, dtype=torch.float) / 128))
    cos, sin = lib.mrope_table(pos.to(DEV), inv.to(DEV))
    fr = pos.float()[:, :, None] * inv[None, None]
    sel = torch.cat([fr[0, :, :16], fr[1, :, 16:40], fr[2, :, 40:64]], -1)  # mrope_section [16,24,24]
    assert (cos.cpu() - sel.cos()).abs().max() < 2e-4 and (sin.cpu() - sel.sin()).abs().max() < 2e-4
    qkv = rnd(S, 2560).bfloat16()
    dq = qkv.clone().to(DEV)
    rows = 1024
    cache_row = (torch.arange(S) + 100).int()
    kc = torch.zeros(rows, 2, 128, dtype=torch.bfloat16, device=DEV)
    vt = torch.zeros(256, rows, dtype=torch.bfloat16, device=DEV)
    lib.mrope_kv(dq, cos, sin, cache_row.to(DEV), kc, vt, 16, 2)
    c = torch.cat([sel.cos()] * 2, -1)[:, None]
    s = torch.cat([sel.sin()] * 2, -1)[:, None]
    rh = lambda x: torch.cat([-x[..., 64:], x[..., :64]], -1)
    q = qkv[:, :2048].float().view(S, 16, 128)
    k = qkv[:, 2048:2304].float().view(S, 2, 128)
    v = qkv[:, 2304:].float().view(S, 2, 128)
    assert rel_max(dq[:, :2048].view(S, 16, 128), q * c + rh(q) * s) < 4e-3
    assert rel_max(kc[100:100 + S], k * c + rh(k) * s) < 4e-3
    assert torch.equal(vt[:, 100:100 + S].cpu(), v.permute(1, 2, 0).reshape(256, S).bfloat16())
    assert kc[:100].abs().max() == 0 and kc[100 + S:].abs().max() == 0  # nothing outside the assigned cache rows


# ---------------------------------------------------------------------------------------------------- LM-head tail
def test_argmax_logprob_entropy():
    lg = rnd(8, 151936)
    dlg = lg.to(DEV)
    assert torch.equal(lib.argmax_rows(dlg).cpu(), lg.argmax(-1))
    tie = torch.zeros(3, 1000); tie[0, 7] = 1; tie[0, 900] = 1; tie[1, 999] = 2
    assert lib.argmax_rows(tie.to(DEV)).tolist() == [7, 999, 0]          # first index wins ties
    lb = lg.bfloat16()
    assert torch.equal(lib.argmax_rows(lb.to(DEV)).cpu(), lb.float().argmax(-1))
    labels = torch.randint(0, 151936, (8,), generator=torch.Generator().manual_seed(0))
    lp, ent = lib.logprob_entropy(dlg, labels.to(DEV))
    np.testing.assert_allclose(lp.cpu().numpy(), host_ref.log_probs_from_logits(lg, labels).numpy(), atol=2e-5)
    np.testing.assert_allclose(ent.cpu().numpy(), host_ref.entropy_from_logits(lg).numpy(), atol=2e-5)


@pytest.mark.parametrize("M,V,K", [(300, 8192, 2048), (5, 8200, 2048), (129, 151936, 2048), (1, 256, 64)])
def test_fused_lm_head_logprob_entropy(M, V, K):
    """sr_lmhead_logprob_entropy (softmax statistics in the tcgen05 GEMM epilogue, no logits in HBM) against
    log_probs_from_logits / entropy_from_logits (roll/utils/functionals.py:193-205) of the fp32 logits of the same bf16
    operands; ragged last N tile (8200 = 32 * 256 + 8), ragged M tile, single row."""
    h = rnd(M, K, scale=1.0, seed=1).bfloat16()
    w = rnd(V, K, scale=0.05, seed=2).bfloat16()
    labels = torch.randint(0, V, (M,), generator=torch.Generator().manual_seed(3))
    labels[0] = V - 1
    lp, ent = lib.lm_head_logprob_entropy(h.to(DEV), w.to(DEV), labels.to(DEV))
    logits = h.double() @ w.double().t()
    lp_ref = host_ref.log_probs_from_logits(logits.float(), labels)
    ent_ref = host_ref.entropy_from_logits(logits.float())
    np.testing.assert_allclose(lp.cpu().numpy(), lp_ref.numpy(), atol=2e-3)
    np.testing.assert_allclose(ent.cpu().numpy(), ent_ref.numpy(), atol=2e-3)
    # against the unfused device path (same accumulators): fp32 logits GEMM + sr_logprob_entropy
    lg = lib.gemm(h.to(DEV), w.to(DEV), out_dtype=torch.float32, block_n=256)
    lp2, ent2 = lib.logprob_entropy(lg, labels.to(DEV))
    np.testing.assert_allclose(lp.cpu().numpy(), lp2.cpu().numpy(), atol=1e-4)
    np.testing.assert_allclose(ent.cpu().numpy(), ent2.cpu().numpy(), atol=1e-4)
    only_lp, none = lib.lm_head_logprob_entropy(h.to(DEV), w.to(DEV), labels.to(DEV), want_entropy=False)
    assert none is None and torch.equal(only_lp, lp)


def test_sampling_topk_topp_against_oracle():
    """temperature -> top-k -> top-p sampling (reference rollout config T=1, top-p 0.8, top-k 100) with given uniforms:
    the picked token must be the oracle's inverse-CDF token (or, at an fp32 boundary, a neighbour whose CDF interval
    touches u within 1e-5)."""
    g = torch.Generator().manual_seed(0)
    V, rows = 151936, 16
    lg = torch.randn(rows, V, generator=g) * 2.0
    lg[3, 100:110] = lg[3].max() + 0.5          # exact ties inside the top-k: lowest index first
    lg[4] = torch.randn(V, generator=g) * 0.01  # flat row: nucleus keeps many candidates
    u = torch.rand(rows, 3, generator=g)
    u[5, :] = torch.tensor([0.0, 0.999999, 0.5])
    for (T, k, p), col in (((1.0, 100, 0.8), 0), ((0.7, 50, 0.95), 1), ((1.3, 1024, 1.0), 2), ((1.0, 1, 0.8), 0)):
        step = torch.tensor([col], dtype=torch.int32, device=DEV)
        got = lib.sample_topk_topp(lg.to(DEV), u.to(DEV), T, k, p, step=step).cpu()
        for r in range(rows):
            ur = float(u[r, col])
            tok, (lo, hi), kept = host_ref.sample_topk_topp_ref(lg[r], ur, T, k, p)
            if int(got[r]) != tok:  # only legal at an fp32 rounding boundary of the CDF: a neighbouring kept candidate
                kl = kept.tolist()
                assert int(got[r]) in kl, (r, T, k, p)
                assert abs(kl.index(int(got[r])) - kl.index(tok)) == 1 and min(abs(ur - lo), abs(ur - hi)) < 1e-4, (r, T, k, p, int(got[r]), tok)
    # top_k = 1 is greedy
    assert torch.equal(lib.sample_topk_topp(lg.to(DEV), u.to(DEV), 1.0, 1, 1.0).cpu(), lg.argmax(-1))


# ---------------------------------------------------------------------------------------------------- mask epilogue / IoU
@pytest.mark.parametrize("i,n_obj", [(0, 3), (1, 1), (2, 4)])
def test_mask_epilogue_bit_exact_and_iou(i, n_obj):
    ml, ms, gt = O.synth_mask_inputs(i, n_obj)
    ref = host_ref.mask_epilogue(ml, ms)
    got = lib.mask_epilogue(ml.to(DEV), ms.to(DEV))
    assert int((got.cpu().numpy() != ref).sum()) == 0                     # byte-exact mask
    iou, cnt = lib.mask_iou(got[None], torch.from_numpy(gt).to(DEV)[None], 1.0, want_counts=True)
    p, g = ref > 0, gt > 0
    assert cnt.cpu().tolist() == [[int((p & g).sum()), int((p | g).sum())]]  # integer counts exact
    assert abs(iou.item() - host_ref.compute_giou(ref, gt)) < 1e-6         # north_star: IoU within 1e-3


def test_mask_epilogue_edge_cases():
    z = torch.zeros(2, 768, 768, dtype=torch.uint8, device=DEV)
    assert lib.mask_iou(z, z, 1.0).tolist() == [1.0, 1.0]                   # both empty -> 1.0 (infer pipeline)
    assert lib.mask_iou(z, z, 0.0).tolist() == [0.0, 0.0]                   # reward-worker convention
    empty = lib.mask_epilogue(torch.zeros(0, 3, 256, 256, device=DEV), torch.zeros(0, 3, device=DEV))
    assert empty.shape == (768, 768) and empty.sum().item() == 0            # no prompts -> zero mask (seg_strategy.py:40-45)
    ml, ms, _ = O.synth_mask_inputs(5, 2)
    a = lib.mask_epilogue(ml.to(DEV), ms.to(DEV))
    both = lib.mask_epilogue(torch.cat([ml, ml]).to(DEV), torch.cat([ms, ms]).to(DEV))
    assert torch.equal(a, both)                                              # OR is idempotent
    ones = torch.ones(1, 768, 768, dtype=torch.uint8, device=DEV) * 255
    assert lib.mask_iou(ones, ones, 1.0).item() == 1.0


def test_seg_strategy_with_the_reference_predictor_api():
    """seg_strategy.py:26-72 through ``create_strategy('seg_infer')`` with a predictor exposing the SAM2ImagePredictor API
    (``set_image`` / ``predict`` -> (masks, scores, low_res_logits)), as ``sam2_seg_model_provider`` returns
    (model_providers.py:515-562): byte-exact with the reference semantics (oracle pinned on the reference's own segment())."""
    from types import SimpleNamespace
    from PIL import Image
    import torch.nn.functional as F
    from socioreasoner_b200.roll_shim import DataProto, create_strategy

    class FakeSam2Predictor:
        def __init__(self, with_low_res=True):
            self.images, self.calls, self.with_low_res = [], [], with_low_res

        def set_image(self, image):
            assert image.size == (756, 756)
            self.images.append(image)

        def predict(self, point_coords=None, point_labels=None, box=None):
            self.calls.append((point_coords, point_labels, box))
            if box is not None and box[0] < 0:
                raise ValueError("bad box")                           # the reference skips objects whose predict() raises
            i = len(self.calls) - 1
            low = ML[i % ML.shape[0]]
            up = F.interpolate(low[None], size=(756, 756), mode="bilinear", align_corners=False)[0]
            return (up > 0).numpy(), MS[i % ML.shape[0]].numpy(), (low.numpy() if self.with_low_res else None)

    ML, MS, _ = O.synth_mask_inputs(5, 4)
    vps = [dict(box=[10, 20, 300, 400]), dict(point_coords=[[50, 60]], point_labels=[1]), dict(box=[-1, 0, 5, 5]),
           dict(box=[1, 2, 3, 4], point_coords=[[5, 6]], point_labels=[1])]
    img = Image.new("RGB", (1024, 1024))
    for with_low in (True, False):
        pred = FakeSam2Predictor(with_low)
        w = SimpleNamespace(worker_config=SimpleNamespace(strategy_args=SimpleNamespace(strategy_name="seg_infer"), model_args=None))
        st = create_strategy(w).initialize(lambda model_args, is_trainable=False: pred)
        d = DataProto.from_dict(tensors=dict(x=torch.zeros(2, 1)), non_tensors=dict(seg_image=[img, img], visual_prompt=[vps, []]))
        out = st.segment(d)["mask"]
        assert len(pred.images) == 1 and len(pred.calls) == 4                 # empty prompt list: the model is not touched
        assert pred.calls[3] == ([[5, 6]], [1], [1, 2, 3, 4])
        keep = [0, 1, 3]                                                      # the failing object is skipped
        ref = host_ref.mask_epilogue(ML[keep], MS[keep])
        assert out[0].dtype == np.uint8 and np.array_equal(out[0], ref)
        assert out[1].shape == (768, 768) and out[1].max() == 0
