"""ctypes binding of libsocioreasoner_b200.so (the C ABI in include/socioreasoner_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every op below hands raw device pointers and the
current CUDA stream to the hand-written sm_100a kernels.  There is NO fallback: if the shared library is missing
or the device is not a B200 the import / first call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p, POINTER
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SR_LIB_PATH: A/B builds of the same C ABI (tools/); the default is the in-tree library built by csrc/Makefile
LIB_PATH = os.environ.get("SR_LIB_PATH") or os.path.join(_HERE, "libsocioreasoner_b200.so")

EPI_NONE, EPI_RESIDUAL, EPI_SWIGLU, EPI_GELU = 0, 1, 2, 3
DT_BF16, DT_F32 = 0, 1

# every symbol include/socioreasoner_b200.h declares: name -> (restype, argtypes)
_P = c_void_p
SYMBOLS = {
    "sr_last_error": (c_char_p, []),
    "sr_abi_version": (c_int, []),
    "sr_device_check": (c_int, []),
    "sr_set_pdl": (c_int, [c_int]),
    "sr_gemm_bf16": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sr_gemm_skinny_bf16": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P]),
    "sr_gemm_skinny_norm_bf16": (c_int, [_P, c_int64, _P, c_float, _P, c_int64, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P]),
    "sr_gemm_skinny_packed_bf16": (c_int, [_P, c_int64, _P, c_float, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P]),
    "sr_gemm_skinny_splitk_workspace_bytes": (c_int64, [c_int]),
    "sr_gemm_skinny_packed_splitk_bf16": (c_int, [_P, c_int64, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "sr_set_gemv_warps": (c_int, [c_int]),
    "sr_set_gemv_load": (c_int, [c_int]),
    "sr_set_gemv_bulk": (c_int, [c_int]),
    "sr_gemm_skinny_bulk_bf16": (c_int, [_P, c_int64, _P, c_float, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P]),
    "sr_attention_fwd": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, c_int64, c_int64, c_int, c_int, _P, c_int64, c_int,
                                 _P, c_int64, c_int64, _P, c_int, _P, _P, c_float, c_int, _P]),
    "sr_decode_attention_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "sr_decode_attention": (c_int, [_P, c_int64, _P, _P, c_int64, _P, _P, c_int, _P, c_int64, c_int, c_int, c_int, c_float, _P, _P]),
    "sr_rmsnorm_bf16": (c_int, [_P, c_int64, _P, _P, c_int64, c_int, c_int, c_float, _P]),
    "sr_gather_rows": (c_int, [_P, c_int64, c_int, _P, _P, c_int64, c_int, c_int, _P]),
    "sr_embed_scatter": (c_int, [_P, c_int64, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int, _P]),
    "sr_vit_rope_table": (c_int, [_P, _P, _P, _P, _P, c_int, _P]),
    "sr_mrope_table": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sr_vit_rope_split": (c_int, [_P, c_int64, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "sr_mrope_kv": (c_int, [_P, c_int64, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "sr_argmax_rows": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P, _P]),
    "sr_sample_topk_topp": (c_int, [_P, c_int64, c_int, c_int, _P, c_int, _P, c_float, c_int, c_float, _P, _P]),
    "sr_logprob_entropy": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P]),
    "sr_set_attention_lean": (c_int, [c_int]),
    "sr_set_attention_trace": (c_int, [_P]),
    "sr_lmhead_logprob_workspace_bytes": (c_int64, [c_int, c_int]),
    "sr_lmhead_logprob_entropy": (c_int, [_P, c_int64, _P, c_int64, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "sr_decode_advance": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "sr_decode_persistent_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "sr_set_decode_persistent_trace": (c_int, [_P, c_int]),
    "sr_set_decode_persistent_tuning": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "sr_decode_step_persistent": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int64, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P,
                                          c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, c_float, c_float, _P, c_int, _P]),
    "sr_mask_epilogue": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "sr_sam2_preprocess": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "sr_sam2_patch_embed": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "sr_layernorm_bf16": (c_int, [_P, c_int64, c_int, c_int, _P, _P, c_float, _P, c_int64, _P]),
    "sr_pool4_rows": (c_int, [_P, c_int64, _P, c_int, c_int, _P, _P]),
    "sr_upsample2x_add": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "sr_elementwise": (c_int, [c_int, _P, _P, _P, c_int64, _P]),
    "sr_sam2_prompt_embed": (c_int, [_P, _P, c_int, _P, _P, _P, c_int, c_float, _P, _P]),
    "sr_small_attention": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int64, _P]),
    "sr_sam2_upscale": (c_int, [_P, _P, _P, _P, _P, c_float, c_int, c_int, _P, _P]),
    "sr_sam2_mask_logits": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, _P, _P]),
    "sr_preprocess_tile": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, _P, _P, c_int, c_int64, _P,
                                   c_int, _P]),
    "sr_mask_iou": (c_int, [_P, _P, c_int, c_int64, c_float, _P, _P, _P]),
}

_lib: Optional[ctypes.CDLL] = None
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)

# Optional per-launch timing (bench.py roofline): when `timeline` is a list, gemm()/attention() bracket their launch
# with CUDA events on the launching stream and append (kind, tag, work, start_event, end_event).
timeline = None


class _Timed:
    __slots__ = ("kind", "tag", "work", "e0")

    def __init__(self, kind, tag, work):
        self.kind, self.tag, self.work = kind, tag, work

    def __enter__(self):
        if timeline is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        if timeline is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream())
            timeline.append((self.kind, self.tag, self.work, self.e0, e1))
        return False


def load() -> ctypes.CDLL:
    """Load the shared library and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C socioreasoner_b200/csrc).  There is no CPU / eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return (load().sr_last_error() or b"").decode()


_SYNC_DEBUG = os.environ.get("SR_SYNC_DEBUG", "0") == "1"


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")
    if _SYNC_DEBUG:  # bring-up aid: surface asynchronous kernel faults at the launch that caused them
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"{what}: kernel fault surfaced at synchronize: {e}") from e


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise RuntimeError(f"{name}: innermost dimension must be contiguous")


def device_check() -> None:
    _check(load().sr_device_check(), "sr_device_check")


# ---------------------------------------------------------------------------------------------------
# ops
# ---------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, epilogue: int = EPI_NONE, out_dtype=torch.bfloat16,
         block_n: int = 0, norm_weight: Optional[torch.Tensor] = None, norm_eps: float = 1e-6) -> torch.Tensor:
    """out[M, n_out] = epilogue(a[M,K] @ w[N,K]^T (+bias)).  a, w bf16 row-major (row stride arbitrary % 8).
    ``norm_weight``: RMSNorm of ``a`` fused into the operand load (skinny/decode shapes only: M <= 32, K % 32 == 0)."""
    global launch_count
    _req(a, torch.bfloat16, "gemm.a"); _req(w, torch.bfloat16, "gemm.w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    assert out.shape == (M, n_out) and out.stride(-1) == 1
    if bias is not None:
        _req(bias, torch.float32, "gemm.bias")
    if residual is not None:
        _req(residual, torch.bfloat16, "gemm.residual")
        epilogue = EPI_RESIDUAL if epilogue == EPI_NONE else epilogue
    if norm_weight is not None:
        _req(norm_weight, torch.float32, "gemm.norm_weight")
        with _Timed("gemm", (M, N, K, epilogue), 2.0 * M * N * K):
            rc = load().sr_gemm_skinny_norm_bf16(a.data_ptr(), a.stride(0), norm_weight.data_ptr(), float(norm_eps), w.data_ptr(),
                                                 w.stride(0), out.data_ptr(), out.stride(0), _ptr(bias), _ptr(residual),
                                                 residual.stride(0) if residual is not None else 0, M, N, K, epilogue,
                                                 DT_F32 if out.dtype == torch.float32 else DT_BF16, _stream())
        _check(rc, "sr_gemm_skinny_norm_bf16")
        launch_count += 1
        return out
    with _Timed("gemm", (M, N, K, epilogue), 2.0 * M * N * K):
        rc = load().sr_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
                                 _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0, M, N, K,
                                 epilogue, DT_F32 if out.dtype == torch.float32 else DT_BF16, block_n, _stream())
    _check(rc, "sr_gemm_bf16")
    launch_count += 1
    return out


def pack_skinny(w: torch.Tensor) -> torch.Tensor:
    """[N,K] -> streaming layout [N/16][K/32][16][32] for ``gemm_skinny_packed`` (pure re-layout, done once at load)."""
    N, K = w.shape
    assert N % 16 == 0 and K % 32 == 0
    return w.view(N // 16, 16, K // 32, 32).permute(0, 2, 1, 3).contiguous()


_splitk_ws = {}


def gemm_skinny_packed(a: torch.Tensor, wp: torch.Tensor, N: int, K: int, out: Optional[torch.Tensor] = None,
                       bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, epilogue: int = EPI_NONE,
                       out_dtype=torch.bfloat16, norm_weight: Optional[torch.Tensor] = None, norm_eps: float = 1e-6,
                       split_k: bool = False) -> torch.Tensor:
    """Decode-step linear on pre-packed weights (see ``pack_skinny``); M = a.shape[0] <= 32.
    ``split_k``: two CTAs per row group split K (long-K, narrow-N layers; M <= 8, no fused norm / SwiGLU)."""
    global launch_count
    _req(a, torch.bfloat16, "gemm_skinny_packed.a"); _req(wp, torch.bfloat16, "gemm_skinny_packed.w")
    M = a.shape[0]
    assert a.shape[1] == K and wp.numel() == N * K and wp.is_contiguous()
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    if residual is not None:
        epilogue = EPI_RESIDUAL if epilogue == EPI_NONE else epilogue
    if split_k and M <= 8 and norm_weight is None and epilogue != EPI_SWIGLU:
        key = (a.device, N)
        ws = _splitk_ws.get(key)
        if ws is None:  # zeroed once (tickets); re-armed by the kernel
            ws = _splitk_ws[key] = torch.zeros((load().sr_gemm_skinny_splitk_workspace_bytes(N),), dtype=torch.uint8, device=a.device)
        with _Timed("gemm", (M, N, K, epilogue), 2.0 * M * N * K):
            rc = load().sr_gemm_skinny_packed_splitk_bf16(a.data_ptr(), a.stride(0), wp.data_ptr(), out.data_ptr(), out.stride(0),
                                                          _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0,
                                                          M, N, K, epilogue, DT_F32 if out.dtype == torch.float32 else DT_BF16,
                                                          ws.data_ptr(), _stream())
        _check(rc, "sr_gemm_skinny_packed_splitk_bf16")
        launch_count += 1
        return out
    with _Timed("gemm", (M, N, K, epilogue), 2.0 * M * N * K):
        rc = load().sr_gemm_skinny_packed_bf16(a.data_ptr(), a.stride(0), _ptr(norm_weight), float(norm_eps), wp.data_ptr(),
                                               out.data_ptr(), out.stride(0), _ptr(bias), _ptr(residual),
                                               residual.stride(0) if residual is not None else 0, M, N, K, epilogue,
                                               DT_F32 if out.dtype == torch.float32 else DT_BF16, _stream())
    _check(rc, "sr_gemm_skinny_packed_bf16")
    launch_count += 1
    return out


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, out: torch.Tensor, tiles: torch.Tensor,
              row_lo: torch.Tensor, row_hi: torch.Tensor, scale: float, head_dim_v: int) -> torch.Tensor:
    """q [Rq,Hq,128], k [Rk,Hkv,128] (views with arbitrary row/head strides), vt [Hkv*hdv, cols], out [Rq,Hq,hdv]."""
    global launch_count
    for t, n in ((q, "q"), (k, "k"), (vt, "vt"), (out, "out")):
        _req(t, torch.bfloat16, "attention." + n)
    _req(tiles, torch.int32, "attention.tiles"); _req(row_lo, torch.int32, "row_lo"); _req(row_hi, torch.int32, "row_hi")
    assert q.shape[2] == 128 and k.shape[2] == 128 and tiles.dim() == 2 and tiles.shape[1] == 4 and tiles.is_contiguous()
    with _Timed("attention", (q.shape[0], q.shape[1], head_dim_v, tiles.shape[0]), 0.0):
        rc = load().sr_attention_fwd(q.data_ptr(), q.stride(0), q.stride(1), q.shape[0], q.shape[1],
                                     k.data_ptr(), k.stride(0), k.stride(1), k.shape[0], k.shape[1],
                                     vt.data_ptr(), vt.stride(0), vt.shape[1], out.data_ptr(), out.stride(0), out.stride(1),
                                     tiles.data_ptr(), tiles.shape[0], row_lo.data_ptr(), row_hi.data_ptr(), float(scale),
                                     head_dim_v, _stream())
    _check(rc, "sr_attention_fwd")
    launch_count += 1
    return out


_dattn_ws = {}


def decode_attention(q, kc, vt, kv_start, kv_len, max_len, out, n_q_heads, n_kv_heads, scale):
    global launch_count
    key = (q.device, q.shape[0], n_kv_heads, int(max_len))
    ws = _dattn_ws.get(key)
    if ws is None:  # zero-initialised once: the leading int32 tickets must start at 0, the kernel re-arms them
        nbytes = load().sr_decode_attention_workspace_bytes(q.shape[0], n_kv_heads, int(max_len))
        ws = _dattn_ws[key] = torch.zeros((nbytes,), dtype=torch.uint8, device=q.device)
    rc = load().sr_decode_attention(q.data_ptr(), q.stride(0), kc.data_ptr(), vt.data_ptr(), vt.stride(0),
                                    kv_start.data_ptr(), kv_len.data_ptr(), int(max_len), out.data_ptr(), out.stride(0),
                                    q.shape[0], n_q_heads, n_kv_heads, float(scale), ws.data_ptr(), _stream())
    _check(rc, "sr_decode_attention")
    launch_count += 1
    return out


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launch_count
    _req(x, torch.bfloat16, "rmsnorm.x"); _req(weight, torch.float32, "rmsnorm.weight")
    rows, dim = x.shape
    if out is None:
        out = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device)
    rc = load().sr_rmsnorm_bf16(x.data_ptr(), x.stride(0), weight.data_ptr(), out.data_ptr(), out.stride(0), rows, dim,
                                float(eps), _stream())
    _check(rc, "sr_rmsnorm_bf16")
    launch_count += 1
    return out


def gather_rows(src: torch.Tensor, idx: Optional[torch.Tensor], out: Optional[torch.Tensor] = None, rows: Optional[int] = None):
    global launch_count
    assert src.dtype in (torch.bfloat16, torch.float32) and src.is_cuda
    n = idx.shape[0] if idx is not None else (rows if rows is not None else src.shape[0])
    if idx is not None:
        _req(idx, torch.int32, "gather.idx")
    if out is None:
        out = torch.empty((n, src.shape[1]), dtype=torch.bfloat16, device=src.device)
    rc = load().sr_gather_rows(src.data_ptr(), src.stride(0), DT_F32 if src.dtype == torch.float32 else DT_BF16, _ptr(idx),
                               out.data_ptr(), out.stride(0), n, src.shape[1], _stream())
    _check(rc, "sr_gather_rows")
    launch_count += 1
    return out


def embed_scatter(table, ids, vis, vis_idx, out=None):
    global launch_count
    _req(table, torch.bfloat16, "embed.table"); _req(ids, torch.int64, "embed.ids")
    rows, cols = ids.shape[0], table.shape[1]
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.bfloat16, device=table.device)
    rc = load().sr_embed_scatter(table.data_ptr(), table.stride(0), ids.data_ptr(), _ptr(vis),
                                 vis.stride(0) if vis is not None else 0, _ptr(vis_idx), out.data_ptr(), out.stride(0),
                                 rows, cols, _stream())
    _check(rc, "sr_embed_scatter")
    launch_count += 1
    return out


def vit_rope_table(hpos, wpos, inv_freq):
    global launch_count
    n = hpos.shape[0]
    cos = torch.empty((n, 40), dtype=torch.float32, device=hpos.device)
    sin = torch.empty_like(cos)
    _check(load().sr_vit_rope_table(hpos.data_ptr(), wpos.data_ptr(), inv_freq.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                    n, _stream()), "sr_vit_rope_table")
    launch_count += 1
    return cos, sin


def mrope_table(pos3: torch.Tensor, inv_freq: torch.Tensor, sections=(16, 24, 24)):
    global launch_count
    _req(pos3, torch.int64, "mrope.pos3")
    assert pos3.dim() == 2 and pos3.shape[0] == 3 and pos3.is_contiguous()
    S = pos3.shape[1]
    cos = torch.empty((S, 64), dtype=torch.float32, device=pos3.device)
    sin = torch.empty_like(cos)
    _check(load().sr_mrope_table(pos3.data_ptr(), inv_freq.data_ptr(), cos.data_ptr(), sin.data_ptr(), S, *sections,
                                 _stream()), "sr_mrope_table")
    launch_count += 1
    return cos, sin


def vit_rope_split(qkv, cos, sin, qp, kp, vt):
    global launch_count
    n = qkv.shape[0]
    _check(load().sr_vit_rope_split(qkv.data_ptr(), qkv.stride(0), cos.data_ptr(), sin.data_ptr(), qp.data_ptr(),
                                    kp.data_ptr(), vt.data_ptr(), vt.stride(0), n, _stream()), "sr_vit_rope_split")
    launch_count += 1


def mrope_kv(qkv, cos, sin, cache_row, kc, vt, n_q_heads, n_kv_heads):
    global launch_count
    _req(cache_row, torch.int32, "mrope_kv.cache_row")
    _check(load().sr_mrope_kv(qkv.data_ptr(), qkv.stride(0), cos.data_ptr(), sin.data_ptr(), cache_row.data_ptr(),
                              kc.data_ptr(), vt.data_ptr(), vt.stride(0), qkv.shape[0], n_q_heads, n_kv_heads, _stream()),
           "sr_mrope_kv")
    launch_count += 1


_argmax_ws = {}


def argmax_rows(logits: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launch_count
    rows, V = logits.shape
    if out is None:
        out = torch.empty((rows,), dtype=torch.int64, device=logits.device)
    key = (logits.device, rows)
    ws = _argmax_ws.get(key)
    if ws is None:
        ws = _argmax_ws[key] = torch.empty((rows * 512,), dtype=torch.uint8, device=logits.device)
    _check(load().sr_argmax_rows(logits.data_ptr(), logits.stride(0), DT_F32 if logits.dtype == torch.float32 else DT_BF16,
                                 rows, V, out.data_ptr(), ws.data_ptr(), _stream()), "sr_argmax_rows")
    launch_count += 2
    return out


def decode_advance(nxt, tokens, step, row0, len0, pos0, cache_row, kv_len, pos3):
    global launch_count
    _check(load().sr_decode_advance(nxt.data_ptr(), tokens.data_ptr(), tokens.shape[1], step.data_ptr(), row0.data_ptr(),
                                    len0.data_ptr(), pos0.data_ptr(), cache_row.data_ptr(), kv_len.data_ptr(), pos3.data_ptr(),
                                    tokens.shape[0], _stream()), "sr_decode_advance")
    launch_count += 1


def decode_layer_table(layers, device) -> torch.Tensor:
    """Device array of ``sr_decode_layer`` structs (8 pointers per layer) for ``decode_step_persistent``."""
    rows = []
    for L in layers:
        assert L.qkv_wp is not None and L.o_wp is not None and L.gu_wp is not None and L.down_wk is not None
        rows.append([L.qkv_wp.data_ptr(), L.qkv_b.data_ptr() if L.qkv_b is not None else 0, L.ln1.data_ptr(), L.o_wp.data_ptr(),
                     L.ln2.data_ptr(), L.gu_wp.data_ptr(), L.down_wk.data_ptr(), 0])
    return torch.tensor(rows, dtype=torch.int64).to(device)


def decode_persistent_workspace(n_q_heads: int, n_kv_heads: int, intermediate: int, device) -> torch.Tensor:
    """Zeroed workspace of ``decode_step_persistent`` (generation counter + flag-carrying exchange buffers, ~5 MB)."""
    return torch.zeros((load().sr_decode_persistent_workspace_bytes(n_q_heads, n_kv_heads, intermediate),), dtype=torch.uint8,
                       device=device)


def decode_persistent_tuning(flags: int = -1, ring_stages: int = 0, offsets=None) -> None:
    """flags bit 0 = evict_first hint on the weight stream, bit 1 = dump plain intermediates (x / qkv / ao / act) for tests,
    bit 4 = K/V L2 prefetch (default 17); ring_stages = shared-memory ring depth; offsets = unit -> CTA rotation per matrix."""
    o = offsets if offsets is not None else (-1, -1, -1, -1)
    _check(load().sr_set_decode_persistent_tuning(int(flags), int(ring_stages), *[int(v) for v in o]), "sr_set_decode_persistent_tuning")


def decode_step_persistent(layer_table: torch.Tensor, lm_head_p: torch.Tensor, final_norm: torch.Tensor, vocab: int,
                           x: torch.Tensor, qkv: torch.Tensor, ao: torch.Tensor, act: torch.Tensor, logits: torch.Tensor,
                           cos: torch.Tensor, sin: torch.Tensor, cache_row: torch.Tensor, kv_start: torch.Tensor,
                           kv_len: torch.Tensor, kc: torch.Tensor, vt: torch.Tensor, n_q_heads: int, n_kv_heads: int,
                           intermediate: int, eps: float, scale: float, workspace: torch.Tensor, stop_after: int = 0,
                           kv_tokens: Optional[int] = None) -> torch.Tensor:
    """One whole decode step (all layers + LM head) in one persistent launch.  kc [L, rows, n_kv, 128], vt [L, n_kv*128, rows];
    x [B, 2048] = the embedded input tokens (read only); logits fp32 [B, vocab].  qkv / ao / act may be None: they (and x) only
    receive plain dumps of the intermediates after ``decode_persistent_tuning(flags=3)`` (tests / bring-up).  ``kv_tokens`` (sum of
    the live sequences' cached tokens, host-side) only feeds the bench timeline: algorithmic bytes = every weight once + the K and
    V rows of the live sequences."""
    global launch_count
    _req(x, torch.bfloat16, "decode_persistent.x"); _req(logits, torch.float32, "decode_persistent.logits")
    _req(kc, torch.bfloat16, "decode_persistent.kc"); _req(vt, torch.bfloat16, "decode_persistent.vt")
    _req(cache_row, torch.int32, "decode_persistent.cache_row"); _req(kv_start, torch.int32, "decode_persistent.kv_start")
    _req(kv_len, torch.int32, "decode_persistent.kv_len"); _req(cos, torch.float32, "decode_persistent.cos")
    B = x.shape[0]
    assert kc.dim() == 4 and vt.dim() == 3 and kc.is_contiguous() and vt.is_contiguous() and layer_table.shape[1] == 8
    assert cos.is_contiguous() and sin.is_contiguous()
    if qkv is not None:
        assert qkv.is_contiguous() and ao.is_contiguous() and act.is_contiguous()
        assert qkv.shape == (B, (n_q_heads + 2 * n_kv_heads) * 128) and ao.shape == (B, n_q_heads * 128) and act.shape == (B, intermediate)
    n_layers, H = int(layer_table.shape[0]), n_q_heads * 128
    work = 2.0 * (n_layers * ((n_q_heads + 2 * n_kv_heads) * 128 * H + H * H + 3 * intermediate * H) + vocab * H)
    work += 2.0 * n_layers * (kv_tokens if kv_tokens is not None else 0) * 2 * n_kv_heads * 128
    with _Timed("decode", (B, n_layers, vocab), work):
        rc = _decode_persistent_call(layer_table, lm_head_p, final_norm, vocab, x, qkv, ao, act, logits, cos, sin, cache_row, kv_start,
                                     kv_len, kc, vt, B, n_q_heads, n_kv_heads, intermediate, eps, scale, workspace, stop_after)
    _check(rc, "sr_decode_step_persistent")
    launch_count += 1
    return logits


def _decode_persistent_call(layer_table, lm_head_p, final_norm, vocab, x, qkv, ao, act, logits, cos, sin, cache_row, kv_start, kv_len,
                            kc, vt, B, n_q_heads, n_kv_heads, intermediate, eps, scale, workspace, stop_after) -> int:
    return load().sr_decode_step_persistent(
        layer_table.data_ptr(), layer_table.shape[0], lm_head_p.data_ptr(), final_norm.data_ptr(), vocab, x.data_ptr(), x.stride(0),
        _ptr(qkv), _ptr(ao), _ptr(act), logits.data_ptr(), logits.stride(0), cos.data_ptr(), sin.data_ptr(),
        cache_row.data_ptr(), kv_start.data_ptr(), kv_len.data_ptr(), kc.data_ptr(), kc.stride(0), vt.data_ptr(), vt.stride(0),
        vt.stride(1), B, n_q_heads, n_kv_heads, intermediate, float(eps), float(scale), workspace.data_ptr(), int(stop_after), _stream())


def sample_topk_topp(logits: torch.Tensor, uniforms: torch.Tensor, temperature: float, top_k: int, top_p: float,
                     step: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """logits fp32 [rows, V]; uniforms fp32 [rows, T] (column *step, or 0) -> sampled token ids [rows]."""
    global launch_count
    _req(logits, torch.float32, "sample.logits"); _req(uniforms, torch.float32, "sample.uniforms")
    rows, V = logits.shape
    if out is None:
        out = torch.empty((rows,), dtype=torch.int64, device=logits.device)
    u2 = uniforms if uniforms.dim() == 2 else uniforms.view(rows, 1)
    _check(load().sr_sample_topk_topp(logits.data_ptr(), logits.stride(0), rows, V, u2.data_ptr(), u2.stride(0), _ptr(step),
                                      float(temperature), int(top_k), float(top_p), out.data_ptr(), _stream()), "sr_sample_topk_topp")
    launch_count += 1
    return out


def logprob_entropy(logits: torch.Tensor, labels: Optional[torch.Tensor], want_entropy: bool = True):
    global launch_count
    rows, V = logits.shape
    lp = torch.empty((rows,), dtype=torch.float32, device=logits.device) if labels is not None else None
    ent = torch.empty((rows,), dtype=torch.float32, device=logits.device) if want_entropy else None
    _check(load().sr_logprob_entropy(logits.data_ptr(), logits.stride(0), DT_F32 if logits.dtype == torch.float32 else DT_BF16,
                                     rows, V, _ptr(labels), _ptr(lp), _ptr(ent), _stream()), "sr_logprob_entropy")
    launch_count += 1
    return lp, ent


def lm_head_logprob_entropy(h: torch.Tensor, w: torch.Tensor, labels: torch.Tensor, want_logprob: bool = True,
                            want_entropy: bool = True):
    """Fused LM head + log-prob / entropy (sr_lmhead_logprob_entropy): h bf16 [M, K] (final-normed hidden rows), w bf16 [V, K]
    (lm_head.weight), labels int64 [M] -> (logprob fp32 [M] | None, entropy fp32 [M] | None).  The logits are never written."""
    global launch_count
    _req(h, torch.bfloat16, "lm_head_logprob.h"); _req(w, torch.bfloat16, "lm_head_logprob.w"); _req(labels, torch.int64, "lm_head_logprob.labels")
    M, K = h.shape
    V = w.shape[0]
    assert w.shape[1] == K and labels.shape == (M,) and h.stride(1) == 1 and w.stride(1) == 1 and labels.is_contiguous()
    lp = torch.empty((M,), dtype=torch.float32, device=h.device) if want_logprob else None
    ent = torch.empty((M,), dtype=torch.float32, device=h.device) if want_entropy else None
    if M == 0:
        return lp, ent
    ws = torch.empty((int(load().sr_lmhead_logprob_workspace_bytes(M, V)),), dtype=torch.uint8, device=h.device)
    with _Timed("gemm", (M, V, K, 4), 2.0 * M * V * K):
        rc = load().sr_lmhead_logprob_entropy(h.data_ptr(), h.stride(0), w.data_ptr(), w.stride(0), labels.data_ptr(), M, V, K,
                                              ws.data_ptr(), _ptr(lp), _ptr(ent), _stream())
    _check(rc, "sr_lmhead_logprob_entropy")
    launch_count += 2
    return lp, ent


def preprocess_tile(image_hwc_u8: torch.Tensor, out_h: int, out_w: int, h_tab, v_tab, scratch: Optional[torch.Tensor], mean3, std3,
                    out: torch.Tensor, row_map: Optional[torch.Tensor] = None, patch_size: int = 14) -> torch.Tensor:
    """uint8 [H, W, 3] device tile -> rows of ``out`` ([N, 6 * patch^2], fp32 or bf16).  h_tab / v_tab = (bounds, coeffs) device int32
    tensors (None when that axis keeps its size); mean3 / std3 = host float32[3] (``preprocess.normalize_constants``)."""
    global launch_count
    _req(image_hwc_u8, torch.uint8, "preprocess.image")
    H, W, C = image_hwc_u8.shape
    assert C == 3 and image_hwc_u8.is_contiguous() and out.stride(1) == 1 and out.dtype in (torch.float32, torch.bfloat16)
    hb, hk = h_tab if h_tab is not None else (None, None)
    vb, vk = v_tab if v_tab is not None else (None, None)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean3])
    d = (ctypes.c_float * 3)(*[float(v) for v in std3])
    _check(load().sr_preprocess_tile(image_hwc_u8.data_ptr(), H, W, out_h, out_w, _ptr(hb), _ptr(hk), hk.shape[1] if hk is not None else 0,
                                     _ptr(vb), _ptr(vk), vk.shape[1] if vk is not None else 0, _ptr(scratch), ctypes.addressof(m),
                                     ctypes.addressof(d), out.data_ptr(), 1 if out.dtype == torch.float32 else 0,
                                     out.stride(0), _ptr(row_map), patch_size, _stream()), "sr_preprocess_tile")
    launch_count += 1 + (1 if hk is not None else 0)
    return out


def mask_epilogue(low_res_logits: torch.Tensor, scores: torch.Tensor, image_size: int = 756, out_size: int = 768,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launch_count
    n_obj = low_res_logits.shape[0]
    if out is None:
        out = torch.empty((out_size, out_size), dtype=torch.uint8, device=low_res_logits.device)
    h, w = (low_res_logits.shape[2], low_res_logits.shape[3]) if n_obj else (1, 1)
    _check(load().sr_mask_epilogue(low_res_logits.data_ptr() if n_obj else None, scores.data_ptr() if n_obj else None, n_obj,
                                   h, w, image_size, out_size, out.data_ptr(), _stream()), "sr_mask_epilogue")
    launch_count += 1
    return out


def mask_iou(pred: torch.Tensor, gt: torch.Tensor, empty_value: float = 1.0, want_counts: bool = False):
    global launch_count
    n = pred.shape[0]
    n_pix = pred[0].numel()
    iou = torch.empty((n,), dtype=torch.float32, device=pred.device)
    counts = torch.empty((n, 2), dtype=torch.int32, device=pred.device) if want_counts else None
    _check(load().sr_mask_iou(pred.data_ptr(), gt.data_ptr(), n, n_pix, float(empty_value), iou.data_ptr(), _ptr(counts),
                              _stream()), "sr_mask_iou")
    launch_count += 1
    return (iou, counts) if want_counts else iou
