"""Index planning for the Hiera trunk (host side, numpy): token orders, q-pooling sources, attention work lists.

Hiera alternates window attention (window_partition / window_unpartition around every block, HF modeling_sam2.py
``Sam2MultiScaleBlock``) with 2x2 max-pooling of the queries at stage changes.  Here every stage keeps its tokens in WINDOW-MAJOR
order for the whole stage (the permutation commutes with every per-row op), so that a window is a contiguous row range and the
partition / unpartition copies disappear; the permutation is applied once by the patch-embedding kernel, once inside stage 3
(whose first block pools out of 4x4 windows into 2x2 groups while its own windows are 16x16) and inverted once per stage output."""
from __future__ import annotations

import numpy as np

from ..plan import AttnPlan


def wm_perm(res: int, ws: int) -> np.ndarray:
    """perm[j] = row-major index (y * res + x) of the token at window-major position j (windows of ws x ws, row-major inside)."""
    j = np.arange(res * res)
    w, i = j // (ws * ws), j % (ws * ws)
    wy, wx = w // (res // ws), w % (res // ws)
    y, x = wy * ws + i // ws, wx * ws + i % ws
    return (y * res + x).astype(np.int32)


def inverse(perm: np.ndarray) -> np.ndarray:
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm), dtype=perm.dtype)
    return inv


def pool_sources(res_in: int, ws_in: int) -> np.ndarray:
    """2x2 max-pool out of window-major(ws_in) tokens at resolution res_in: idx[j, 0..3] = input positions of the four sources of
    output position j, outputs in window-major(ws_in / 2) order at resolution res_in / 2."""
    res_out, ws_out = res_in // 2, ws_in // 2
    rm_out = wm_perm(res_out, ws_out)
    inv_in = inverse(wm_perm(res_in, ws_in))
    y, x = rm_out // res_out, rm_out % res_out
    idx = np.stack([inv_in[(2 * y + dy) * res_in + (2 * x + dx)] for dy in (0, 1) for dx in (0, 1)], 1)
    return idx.astype(np.int32)


def window_plan(n_windows: int, q_per_win: int, kv_per_win: int) -> AttnPlan:
    """Attention work list for window-major tokens: query rows of window w = [w*q_per_win, ...), keys [w*kv_per_win, ...).
    Windows with fewer than 128 query rows are packed several per 128-row tile (each row keeps its own key range)."""
    tiles = []
    n_q = n_windows * q_per_win
    row_lo = (np.arange(n_q) // q_per_win) * kv_per_win
    row_hi = row_lo + kv_per_win
    if q_per_win >= 128:
        for w in range(n_windows):
            for r0 in range(0, q_per_win, 128):
                tiles.append((w * q_per_win + r0, min(128, q_per_win - r0), w * kv_per_win, (w + 1) * kv_per_win))
    else:
        per = 128 // q_per_win
        for w0 in range(0, n_windows, per):
            w1 = min(n_windows, w0 + per)
            tiles.append((w0 * q_per_win, (w1 - w0) * q_per_win, w0 * kv_per_win, w1 * kv_per_win))
    return AttnPlan(np.asarray(tiles, dtype=np.int32), row_lo.astype(np.int32), row_hi.astype(np.int32))


def full_plan(n_q: int, n_kv: int) -> AttnPlan:
    tiles = [(r0, min(128, n_q - r0), 0, n_kv) for r0 in range(0, n_q, 128)]
    return AttnPlan(np.asarray(tiles, dtype=np.int32), np.zeros(n_q, np.int32), np.full(n_q, n_kv, np.int32))
