"""Stand-alone timing of the tcgen05 GEMM on the large-M shapes of the bench step (CUDA events, back-to-back launches of one
shape, correctness spot check against torch fp32 on a row sample).  SR_LIB_PATH selects an alternative build of the same C ABI.
Usage: python tools/gemm_bench.py [iters=30] [forced block_n list, e.g. 256,160,128]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
BNS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else []   # also time these forced tile widths
NONE, RES, SWIGLU = lib.EPI_NONE, lib.EPI_RESIDUAL, lib.EPI_SWIGLU
SHAPES = [  # (name, M, N, K, epilogue, bias)
    ("vit qkv", 23328, 3840, 1280, NONE, True), ("vit proj", 23328, 1280, 1280, RES, True),
    ("vit fc1 (SwiGLU)", 23328, 6912, 1280, SWIGLU, True), ("vit fc2", 23328, 1280, 3456, RES, True),
    ("dec qkv", 6936, 2560, 2048, NONE, True), ("dec o", 6936, 2048, 2048, RES, False),
    ("dec gate/up (SwiGLU)", 6936, 22016, 2048, SWIGLU, False), ("dec down", 6936, 2048, 11008, RES, False),
]
dev = torch.device("cuda")
print("lib:", lib.LIB_PATH)
g = torch.Generator(device=dev).manual_seed(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot_f = tot_ms = 0.0
for name, M, N, K, epi, has_bias in SHAPES:
    a = (torch.randn((M, K), device=dev, generator=g) * 0.5).bfloat16()
    w = (torch.randn((N, K), device=dev, generator=g) * (K ** -0.5)).bfloat16()
    bias = torch.randn((N,), device=dev, generator=g) if has_bias else None
    n_out = N // 2 if epi == SWIGLU else N
    res = (torch.randn((M, n_out), device=dev, generator=g)).bfloat16() if epi == RES else None
    out = torch.empty((M, n_out), dtype=torch.bfloat16, device=dev)
    alt = []
    for bn in BNS:
        for _ in range(3):
            lib.gemm(a, w, out=out, bias=bias, residual=res, epilogue=epi, block_n=bn)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            lib.gemm(a, w, out=out, bias=bias, residual=res, epilogue=epi, block_n=bn)
        e1.record(); torch.cuda.synchronize()
        alt.append(f"bn{bn}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us")
    for _ in range(3):
        lib.gemm(a, w, out=out, bias=bias, residual=res, epilogue=epi)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        lib.gemm(a, w, out=out, bias=bias, residual=res, epilogue=epi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # correctness on a row sample (first / middle / last tile rows)
    rows = torch.tensor([0, 1, 127, 128, M // 2, M - 129, M - 1], device=dev)
    y = a[rows].float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if epi == SWIGLU:  # packed rows: [16 gate | 16 up] per 32 columns
        y = y.view(len(rows), N // 32, 2, 16)
        y = (torch.nn.functional.silu(y[:, :, 0]) * y[:, :, 1]).reshape(len(rows), N // 2)
    if res is not None:
        y = y + res[rows].float()
    err = ((out[rows].float() - y).norm() / y.norm()).item()
    fl = 2.0 * M * N * K
    tot_f += fl; tot_ms += ms
    print(f"{name:22s} M{M} N{N} K{K}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  rel err {err:.2e}  {' '.join(alt)}", flush=True)
print(f"sum over shapes: {tot_f / tot_ms / 1e9:.1f} TFLOP/s")
