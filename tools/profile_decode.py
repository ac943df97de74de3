"""The kernels of ONE decode layer at bench sizes (B=8, ~900 cached tokens), cold weights, for `ncu --set full`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from socioreasoner_b200 import lib
dev = "cuda"
torch.manual_seed(0)
B, L = 8, 1024
layers = 6  # distinct weight sets so that every launch streams cold weights
def w(n, k): return (torch.randn(n, k, device=dev) * 0.02).bfloat16()
W = [dict(qkv=w(2560, 2048), o=w(2048, 2048), gu=w(22016, 2048), down=w(2048, 11008), ln=torch.ones(2048, device=dev),
          b=torch.zeros(2560, device=dev)) for _ in range(layers)]
for l in W:
    for k in ("qkv", "o", "gu", "down"):
        l[k + "p"] = lib.pack_skinny(l[k])
x = torch.randn(B, 2048, device=dev).bfloat16()
qkv = torch.empty(B, 2560, device=dev, dtype=torch.bfloat16)
ao = torch.randn(B, 2048, device=dev).bfloat16()
act = torch.empty(B, 11008, device=dev, dtype=torch.bfloat16)
kc = torch.randn(B * L, 2, 128, device=dev).bfloat16()
vt = torch.randn(256, B * L, device=dev).bfloat16()
start = (torch.arange(B, device=dev) * L).int()
lens = torch.full((B,), 900, device=dev, dtype=torch.int32)
for rep in range(2):
    for l in W:
        lib.gemm_skinny_packed(x, l["qkvp"], 2560, 2048, out=qkv, bias=l["b"], norm_weight=l["ln"])
        lib.decode_attention(qkv, kc, vt, start, lens, L, ao, 16, 2, 128 ** -0.5)
        lib.gemm_skinny_packed(ao, l["op"], 2048, 2048, out=x, residual=x)
        lib.gemm_skinny_packed(x, l["gup"], 22016, 2048, out=act, epilogue=lib.EPI_SWIGLU, norm_weight=l["ln"])
        lib.gemm_skinny_packed(act, l["downp"], 2048, 11008, out=x, residual=x)
torch.cuda.synchronize()
print("done")
