"""Cold-cache timing of the skinny GEMM: cycles through enough distinct weight copies to defeat the 126 MB L2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib
dev = "cuda"
def run(name, N, K, epi=lib.EPI_NONE, f32=False, norm=False, M=8, packed=False, split_k=False):
    nbytes = N * K * 2
    copies = max(2, int(400e6 // nbytes) + 1)
    Ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(copies)]
    if packed:
        Ws = [lib.pack_skinny(w) for w in Ws]
    x = torch.randn(M, K, device=dev).bfloat16()
    nw = torch.ones(K, device=dev) if norm else None
    n_out = N // 2 if epi == lib.EPI_SWIGLU else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    call = (lambda w: lib.gemm_skinny_packed(x, w, N, K, out=out, epilogue=epi, norm_weight=nw, split_k=split_k)) if packed else \
           (lambda w: lib.gemm(x, w, out=out, epilogue=epi, norm_weight=nw))
    for w in Ws: call(w)
    torch.cuda.synchronize()
    iters = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for w in Ws: call(w)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * copies)
    print(f"{name:28s} N={N:6d} K={K:5d} copies={copies:3d}: {us:7.2f} us  {nbytes / us / 1e6:6.2f} TB/s", flush=True)
print("=== split-K")
run("down", 2048, 11008, packed=True)
run("down split-K", 2048, 11008, packed=True, split_k=True)
run("o", 2048, 2048, packed=True)
run("o split-K", 2048, 2048, packed=True, split_k=True)
run("K8192", 2752, 8192, packed=True)
run("K8192 split-K", 2752, 8192, packed=True, split_k=True)
