"""The product backend of ``Sam2Engine``: every op is one call into ``libsocioreasoner_b200.so`` (tcgen05 GEMM and attention from
the VLM path, the SAM2-specific kernels of ``csrc/sr_sam2.cu``).  torch only allocates buffers and concatenates rows."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import lib


class DevPlan:
    def __init__(self, p, device):
        self.tiles = torch.from_numpy(np.ascontiguousarray(p.tiles)).to(device)
        self.row_lo = torch.from_numpy(np.ascontiguousarray(p.row_lo)).to(device)
        self.row_hi = torch.from_numpy(np.ascontiguousarray(p.row_hi)).to(device)


class CudaOps:
    act_dtype = torch.bfloat16
    supports_graph = True   # every op is a stream-ordered launch on torch's current stream: the engine may capture CUDA graphs

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.L = lib.load()

    def _s(self):
        return lib._stream()

    def _bump(self, n=1):
        lib.launch_count += n

    def make_plan(self, p):
        return DevPlan(p, self.device)

    # ---- tensor-core ops (shared with the VLM path)
    def linear(self, x, w, bias=None, residual=None, act=None, out_dtype=None):
        assert act in (None, "gelu")
        if x.stride(-1) != 1:
            x = x.contiguous()
        return lib.gemm(x, w, bias=bias, residual=residual, epilogue=lib.EPI_GELU if act == "gelu" else lib.EPI_NONE,
                        out_dtype=out_dtype or torch.bfloat16)

    def attention(self, q, k, vt, plan, scale, hdv):
        n, H = q.shape[0], q.shape[1]
        out = torch.empty((n, H, hdv), dtype=torch.bfloat16, device=self.device)
        lib.attention(q, k, vt, out, plan.tiles, plan.row_lo, plan.row_hi, scale, hdv)
        return out.view(n, H * hdv)

    def gather_rows(self, x, idx):
        return lib.gather_rows(x, idx)

    # ---- csrc/sr_sam2.cu
    def preprocess_image(self, arr_hwc_u8: np.ndarray, size: int, mean, std):
        img = torch.from_numpy(np.ascontiguousarray(arr_hwc_u8)).to(self.device, non_blocking=True)
        out = torch.empty((3, size, size), dtype=torch.float32, device=self.device)
        m, d = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
        lib._check(self.L.sr_sam2_preprocess(img.data_ptr(), img.shape[0], img.shape[1], size, ctypes.addressof(m), ctypes.addressof(d),
                                             out.data_ptr(), self._s()), "sr_sam2_preprocess")
        self._bump()
        return out

    def patch_embed(self, pixels, W):
        S = pixels.shape[-1]
        C = W.patch_w.shape[0]
        k = int(round((W.patch_w.shape[1] // 3) ** 0.5))
        stride = S // int(round(W.patch_rows.numel() ** 0.5))
        out = torch.empty((W.patch_rows.numel(), C), dtype=torch.bfloat16, device=self.device)
        px = pixels.to(self.device, torch.float32).contiguous()
        lib._check(self.L.sr_sam2_patch_embed(px.data_ptr(), S, C, k, stride, k // 2, W.patch_w.data_ptr(), W.patch_b.data_ptr(),
                                              W.pos_table.data_ptr(), W.patch_rows.data_ptr(), out.data_ptr(), self._s()), "sr_sam2_patch_embed")
        self._bump()
        return out

    def layernorm(self, x, w, b, eps):
        assert x.dtype == torch.bfloat16 and x.stride(1) == 1
        out = torch.empty((x.shape[0], x.shape[1]), dtype=torch.bfloat16, device=self.device)
        lib._check(self.L.sr_layernorm_bf16(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], w.data_ptr(), b.data_ptr(), float(eps),
                                            out.data_ptr(), out.stride(0), self._s()), "sr_layernorm_bf16")
        self._bump()
        return out

    def pool4(self, x, idx):
        assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and idx.dtype == torch.int32 and idx.is_contiguous()
        out = torch.empty((idx.shape[0], x.shape[1]), dtype=torch.bfloat16, device=self.device)
        lib._check(self.L.sr_pool4_rows(x.data_ptr(), x.stride(0), idx.data_ptr(), idx.shape[0], x.shape[1], out.data_ptr(), self._s()),
                   "sr_pool4_rows")
        self._bump()
        return out

    def upsample2x_add(self, a, b, r):
        out = torch.empty_like(a)
        lib._check(self.L.sr_upsample2x_add(a.data_ptr(), b.data_ptr(), r, a.shape[1], out.data_ptr(), self._s()), "sr_upsample2x_add")
        self._bump()
        return out

    def add(self, a, b):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        lib._check(self.L.sr_elementwise(0, a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), self._s()), "sr_elementwise")
        self._bump()
        return out

    def relu(self, x):
        assert x.is_contiguous() and x.dtype == torch.bfloat16
        lib._check(self.L.sr_elementwise(1, x.data_ptr(), None, None, x.numel(), self._s()), "sr_elementwise")
        self._bump()
        return x

    def sigmoid(self, x):
        assert x.is_contiguous() and x.dtype == torch.float32
        lib._check(self.L.sr_elementwise(2, x.data_ptr(), None, None, x.numel(), self._s()), "sr_elementwise")
        self._bump()
        return x

    def cat_rows(self, *ts):
        return torch.cat([t.to(torch.bfloat16) for t in ts], 0)

    def small_attention(self, q, k, v, heads):
        d = q.shape[1] // heads
        out = torch.empty((q.shape[0], q.shape[1]), dtype=torch.bfloat16, device=self.device)
        lib._check(self.L.sr_small_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), q.shape[0],
                                             k.shape[0], heads, d, out.data_ptr(), out.stride(0), self._s()), "sr_small_attention")
        self._bump()
        return out

    def prompt_embed(self, coords, labels, W, image_size):
        if torch.is_tensor(coords):   # static device buffers of a captured prompt graph
            c, l = coords, labels
            assert c.is_cuda and c.dtype == torch.float32 and l.dtype == torch.int32 and c.is_contiguous()
        else:
            c = torch.as_tensor(np.asarray(coords, np.float32)).to(self.device)
            l = torch.as_tensor(np.asarray(labels, np.int32)).to(self.device)
        hidden = W.point_embed.shape[1]
        out = torch.empty((c.shape[0], hidden), dtype=torch.bfloat16, device=self.device)
        lib._check(self.L.sr_sam2_prompt_embed(c.data_ptr(), l.data_ptr(), c.shape[0], W.pe_gauss.data_ptr(), W.point_embed.data_ptr(),
                                               W.not_a_point.data_ptr(), hidden, float(image_size), out.data_ptr(), self._s()),
                   "sr_sam2_prompt_embed")
        self._bump()
        return out

    def _upscale(self, y, feat, bias, ln_w, ln_b, r):
        C = bias.shape[0]
        out = torch.empty((4 * r * r, C), dtype=torch.bfloat16, device=self.device)
        lib._check(self.L.sr_sam2_upscale(y.data_ptr(), feat.data_ptr(), bias.data_ptr(), lib._ptr(ln_w), lib._ptr(ln_b), 1e-6, r, C,
                                          out.data_ptr(), self._s()), "sr_sam2_upscale")
        self._bump()
        return out

    def upscale1(self, y, feat, bias, ln_w, ln_b, r):
        return self._upscale(y, feat, bias, ln_w, ln_b, r)

    def upscale2(self, y, feat, bias, r):
        return self._upscale(y, feat, bias, None, None, r)

    def mask_logits(self, hyper, up, first, count):
        out = torch.empty((count, up.shape[0]), dtype=torch.float32, device=self.device)
        lib._check(self.L.sr_sam2_mask_logits(hyper.data_ptr(), up.data_ptr(), first, count, up.shape[1], up.shape[0], out.data_ptr(),
                                              self._s()), "sr_sam2_mask_logits")
        self._bump()
        return out
