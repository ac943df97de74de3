"""TEST INFRASTRUCTURE: a plain-torch (CPU, fp32) implementation of the ``ops`` interface ``Sam2Engine`` drives.  It lets the CPU
suite run the engine's host logic -- token orders, pooling tables, attention work lists, padded / folded weights -- end to end
against the HF oracle.  The product has exactly one backend (``socioreasoner_b200.sam2.ops_cuda.CudaOps``)."""
import numpy as np
import torch
import torch.nn.functional as F


class TorchOps:
    act_dtype = torch.float32

    def make_plan(self, p):
        return p

    def linear(self, x, w, bias=None, residual=None, act=None, out_dtype=None):
        y = x.float() @ w.float().t()
        if bias is not None:
            y = y + bias
        if act == "gelu":
            y = F.gelu(y)
        if residual is not None:
            y = y + residual
        return y

    def layernorm(self, x, w, b, eps):
        return F.layer_norm(x.float(), (x.shape[1],), w, b, eps)

    def pool4(self, x, idx):
        return x[idx.long()].amax(1)

    def gather_rows(self, x, idx):
        return x[idx.long()]

    def attention(self, q, k, vt, plan, scale, hdv):
        n, H, _ = q.shape
        out = torch.zeros(n, H, hdv)
        covered = np.zeros(n, bool)
        for q0, nq, lo, hi in plan.tiles.tolist():        # honour the work list exactly as the kernel does
            for r in range(q0, q0 + nq):
                a, b_ = int(plan.row_lo[r]), int(plan.row_hi[r])
                assert lo <= a and b_ <= hi and not covered[r]
                covered[r] = True
        assert covered.all()
        lo, hi = torch.from_numpy(plan.row_lo).long(), torch.from_numpy(plan.row_hi).long()
        # rows with identical ranges are processed together
        keys, inv = torch.unique(torch.stack([lo, hi], 1), dim=0, return_inverse=True)
        for g, (a, b_) in enumerate(keys.tolist()):
            rows = (inv == g).nonzero().flatten()
            s = torch.einsum("nhd,mhd->hnm", q[rows].float(), k[a:b_].float()) * scale
            p = s.softmax(-1)
            v = vt[:, a:b_].view(H, hdv, b_ - a).float()
            out[rows] = torch.einsum("hnm,hdm->nhd", p, v)
        return out.reshape(n, H * hdv)

    def patch_embed(self, pixels, W):
        C0 = W.patch_w.shape[0]
        k = int(round((W.patch_w.shape[1] // 3) ** 0.5))
        y = F.conv2d(pixels[None].float(), W.patch_w.view(C0, 3, k, k), W.patch_b, stride=4, padding=k // 2)[0]   # [C0, R, R]
        rm = y.permute(1, 2, 0).reshape(-1, C0)
        out = torch.empty_like(rm)
        out[W.patch_rows.long()] = rm
        return out + W.pos_table

    def upsample2x_add(self, a, b, r):
        C = a.shape[1]
        up = b.view(r, r, C).repeat_interleave(2, 0).repeat_interleave(2, 1).reshape(4 * r * r, C)
        return a + up

    def add(self, a, b):
        return a + b

    def relu(self, x):
        return x.clamp_min(0)

    def sigmoid(self, x):
        return torch.sigmoid(x)

    def cat_rows(self, *ts):
        return torch.cat([t.float() for t in ts], 0)

    def small_attention(self, q, k, v, heads):
        Mq, D = q.shape
        d = D // heads
        s = torch.einsum("qhd,khd->hqk", q.view(Mq, heads, d), k.view(-1, heads, d)) * d ** -0.5
        return torch.einsum("hqk,khd->qhd", s.softmax(-1), v.view(-1, heads, d)).reshape(Mq, D)

    def prompt_embed(self, coords, labels, W, image_size):
        c = (torch.as_tensor(coords, dtype=torch.float32) + 0.5) / image_size
        lab = torch.as_tensor(labels, dtype=torch.long)
        e = (2 * c - 1) @ W.pe_gauss
        e = 2 * np.pi * e
        e = torch.cat([e.sin(), e.cos()], -1)
        e = torch.where(lab[:, None] == -1, W.not_a_point[None], e)
        return e + W.point_embed[lab.clamp(min=0)] * (lab >= 0)[:, None]

    def _shuffle(self, y, c, r):       # [r*r, 4*c] (dy, dx, c) -> [4*r*r, c] row-major at 2r x 2r
        return y.view(r, r, 2, 2, c).permute(0, 2, 1, 3, 4).reshape(4 * r * r, c)

    def upscale1(self, y, feat, bias, ln_w, ln_b, r):
        c = bias.shape[0]
        z = self._shuffle(y.float(), c, r) + bias + feat
        return F.gelu(F.layer_norm(z, (c,), ln_w, ln_b, 1e-6))

    def upscale2(self, y, feat, bias, r):
        c = bias.shape[0]
        return F.gelu(self._shuffle(y.float(), c, r) + bias + feat)

    def mask_logits(self, hyper, up, first, count):
        return hyper[first:first + count].float() @ up.float().t()
