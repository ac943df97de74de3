"""A few launches of the hot GEMM shapes for `ncu --set full` (one GPU, short)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib
dev = "cuda"
torch.manual_seed(0)
def mk(M, N, K):
    return (torch.randn(M, K, device=dev) * 0.5).bfloat16(), (torch.randn(N, K, device=dev) * 0.05).bfloat16()
shapes = {"qkv": (23328, 3840, 1280, lib.EPI_NONE, True, False), "swiglu": (23328, 6912, 1280, lib.EPI_SWIGLU, True, False),
          "proj": (23328, 1280, 1280, lib.EPI_NONE, True, True), "decgu": (6936, 22016, 2048, lib.EPI_SWIGLU, False, False)}
which = sys.argv[1:] or list(shapes)
for name in which:
    M, N, K, epi, bias, res = shapes[name]
    a, w = mk(M, N, K)
    b = torch.randn(N, device=dev) * 0.1 if bias else None
    n_out = N // 2 if epi == lib.EPI_SWIGLU else N
    r = torch.randn(M, n_out, device=dev).bfloat16() if res else None
    out = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        lib.gemm(a, w, out=out, bias=b, residual=r, epilogue=epi)
    torch.cuda.synchronize()
print("done")
