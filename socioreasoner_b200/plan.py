"""Host-side index planning for the Qwen2.5-VL hot path (numpy, no device work).

* ViT token order / window index / 2-D RoPE position ids -- follows HF
  ``transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py:382-451`` (``rot_pos_emb``, ``get_window_index``), which the
  reference instantiates at ``mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py:196-200``.
* Attention work lists for ``sr_attention_fwd``: (tiles, row_lo, row_hi) expressing windowed / full / causal
  attention over packed variable-length sequences (``cu_window_seqlens`` / ``cu_seqlens`` at HF:469-507).
* M-RoPE position ids with the REFERENCE semantics (``mcore_adapter/.../modeling_qwen2_5_vl.py:319-441``; image
  temporal index 0), used when the caller does not pass ``position_ids``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

TILE_ROWS = 128


# ---------------------------------------------------------------------------------------------------
# attention plans
# ---------------------------------------------------------------------------------------------------
@dataclass
class AttnPlan:
    tiles: np.ndarray    # int32 [n_tiles, 4] = q_row0, q_nrows, kv_lo, kv_hi
    row_lo: np.ndarray   # int32 [n_rows]
    row_hi: np.ndarray   # int32 [n_rows]


def plan_from_segments(seg_bounds: np.ndarray, n_rows: int, break_at: Optional[np.ndarray] = None,
                       pack_whole: bool = False) -> AttnPlan:
    """Bidirectional attention inside each segment.  ``seg_bounds`` = sorted boundaries [0, ..., n_rows] (q rows and
    key rows share one index space).  Tiles are 128-row chunks of each span between consecutive ``break_at`` boundaries
    (default: one span = everything).  ``pack_whole``: a tile is a maximal run of WHOLE segments with at most 128 rows
    (short segments: the ViT windows of 36-64 tokens), so that its key range is its own rows -- one 128-key block per tile,
    where blind 128-row chunks straddle three windows and need two."""
    seg_bounds = np.asarray(seg_bounds, dtype=np.int64)
    assert seg_bounds[0] == 0 and seg_bounds[-1] == n_rows and np.all(np.diff(seg_bounds) > 0)
    seg_id = np.searchsorted(seg_bounds, np.arange(n_rows), side="right") - 1
    row_lo = seg_bounds[seg_id].astype(np.int32)
    row_hi = seg_bounds[seg_id + 1].astype(np.int32)
    tiles = []
    if pack_whole:
        i, n_seg = 0, len(seg_bounds) - 1
        while i < n_seg:
            r0 = int(seg_bounds[i])
            if seg_bounds[i + 1] - r0 > TILE_ROWS:          # a long segment: 128-row chunks, all keys of the segment
                for c0 in range(r0, int(seg_bounds[i + 1]), TILE_ROWS):
                    c1 = min(c0 + TILE_ROWS, int(seg_bounds[i + 1]))
                    tiles.append((c0, c1 - c0, r0, int(seg_bounds[i + 1])))
                i += 1
                continue
            j = i + 1
            while j < n_seg and seg_bounds[j + 1] - r0 <= TILE_ROWS:
                j += 1
            r1 = int(seg_bounds[j])
            tiles.append((r0, r1 - r0, r0, r1))
            i = j
        return AttnPlan(np.asarray(tiles, dtype=np.int32).reshape(-1, 4), row_lo, row_hi)
    if break_at is None:
        break_at = np.array([0, n_rows], dtype=np.int64)
    for a, b in zip(break_at[:-1], break_at[1:]):
        for r0 in range(int(a), int(b), TILE_ROWS):
            r1 = min(r0 + TILE_ROWS, int(b))
            tiles.append((r0, r1 - r0, int(row_lo[r0]), int(row_hi[r1 - 1])))
    return AttnPlan(np.asarray(tiles, dtype=np.int32).reshape(-1, 4), row_lo, row_hi)


def plan_causal(lengths: Sequence[int], q_starts: Sequence[int], kv_starts: Sequence[int]) -> AttnPlan:
    """Causal self-attention per sample.  Sample b has ``lengths[b]`` tokens; its query rows are packed at
    ``q_starts[b]`` and its keys live at cache rows ``kv_starts[b] + pos``."""
    n_rows = int(q_starts[-1] + lengths[-1]) if len(lengths) else 0
    row_lo = np.zeros(n_rows, dtype=np.int32)
    row_hi = np.zeros(n_rows, dtype=np.int32)
    tiles = []
    for L, q0, k0 in zip(lengths, q_starts, kv_starts):
        L, q0, k0 = int(L), int(q0), int(k0)
        row_lo[q0:q0 + L] = k0
        row_hi[q0:q0 + L] = k0 + 1 + np.arange(L)
        for r0 in range(0, L, TILE_ROWS):
            r1 = min(r0 + TILE_ROWS, L)
            tiles.append((q0 + r0, r1 - r0, k0, k0 + r1))
    # longest key range first: CTAs are scheduled in tile order, so the 7-block tiles of a long prompt start in the first
    # wave instead of forming the tail of the launch
    tiles.sort(key=lambda t: t[2] - t[3])
    return AttnPlan(np.asarray(tiles, dtype=np.int32).reshape(-1, 4), row_lo, row_hi)


# ---------------------------------------------------------------------------------------------------
# ViT indices
# ---------------------------------------------------------------------------------------------------
@dataclass
class VitPlan:
    n: int                      # total patches
    m: int                      # total merged tokens
    row_perm: np.ndarray        # int32 [n]: window-ordered row i reads processor row row_perm[i]
    merged_reverse: np.ndarray  # int32 [m]: original-order merged token j sits at window-ordered index merged_reverse[j]
    hpos: np.ndarray            # int32 [n] (window order)
    wpos: np.ndarray            # int32 [n]
    cu_window_seqlens: np.ndarray
    cu_seqlens: np.ndarray
    window_plan: AttnPlan
    full_plan: AttnPlan


def vit_plan(grid_thw: np.ndarray, spatial_merge_size: int = 2, window_size: int = 112, patch_size: int = 14) -> VitPlan:
    grid_thw = np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3)
    ms = spatial_merge_size
    unit = ms * ms
    vws = window_size // ms // patch_size  # merged units per window side (4)
    window_index: List[np.ndarray] = []
    cu_win = [0]
    hpos_all, wpos_all = [], []
    win_id0 = 0
    for t, h, w in grid_thw.tolist():
        # rot_pos_emb (HF:382-409): (h, w) ids in merge-block order
        hp = np.broadcast_to(np.arange(h)[:, None], (h, w)).reshape(h // ms, ms, w // ms, ms).transpose(0, 2, 1, 3).reshape(-1)
        wp = np.broadcast_to(np.arange(w)[None, :], (h, w)).reshape(h // ms, ms, w // ms, ms).transpose(0, 2, 1, 3).reshape(-1)
        hpos_all.append(np.tile(hp, t))
        wpos_all.append(np.tile(wp, t))
        # get_window_index (HF:411-451)
        gh, gw = h // ms, w // ms
        index = np.arange(t * gh * gw).reshape(t, gh, gw)
        pad_h = vws - gh % vws
        pad_w = vws - gw % vws
        nwh, nww = (gh + pad_h) // vws, (gw + pad_w) // vws
        padded = np.full((t, gh + pad_h, gw + pad_w), -100, dtype=np.int64)
        padded[:, :gh, :gw] = index
        padded = padded.reshape(t, nwh, vws, nww, vws).transpose(0, 1, 3, 2, 4).reshape(t, nwh * nww, vws, vws)
        seqlens = (padded != -100).sum(axis=(2, 3)).reshape(-1)
        flat = padded.reshape(-1)
        window_index.append(flat[flat != -100] + win_id0)
        cu_win.extend((np.cumsum(seqlens) * unit + cu_win[-1]).tolist())
        win_id0 += t * gh * gw
    window_index = np.concatenate(window_index)
    # unique_consecutive (HF:469-474)
    cu_win = np.asarray(cu_win, dtype=np.int64)
    cu_win = cu_win[np.concatenate([[True], np.diff(cu_win) != 0])]
    n = int((grid_thw[:, 0] * grid_thw[:, 1] * grid_thw[:, 2]).sum())
    m = n // unit
    row_perm = (window_index[:, None] * unit + np.arange(unit)[None, :]).reshape(-1)
    hpos = np.concatenate(hpos_all)[row_perm]
    wpos = np.concatenate(wpos_all)[row_perm]
    per_frame = np.repeat(grid_thw[:, 1] * grid_thw[:, 2], grid_thw[:, 0])
    cu = np.concatenate([[0], np.cumsum(per_frame)])
    return VitPlan(
        n=n, m=m, row_perm=row_perm.astype(np.int32), merged_reverse=np.argsort(window_index).astype(np.int32),
        hpos=hpos.astype(np.int32), wpos=wpos.astype(np.int32), cu_window_seqlens=cu_win, cu_seqlens=cu,
        window_plan=plan_from_segments(cu_win, n, pack_whole=True), full_plan=plan_from_segments(cu, n, break_at=cu),
    )


# ---------------------------------------------------------------------------------------------------
# M-RoPE position ids, reference semantics (mcore_adapter/.../qwen2_5_vl/modeling_qwen2_5_vl.py:319-441)
# ---------------------------------------------------------------------------------------------------
def rope_index(input_ids: np.ndarray, image_grid_thw: Optional[np.ndarray], attention_mask: Optional[np.ndarray],
               image_token_id: int, vision_start_token_id: int, spatial_merge_size: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    """-> (position_ids int64 [3,B,S], rope_deltas int64 [B,1]).  Text: running index on all three axes; an image span
    of merged grid (H, W) starting after ``text_len`` text tokens gets t = base, h = base + row, w = base + col with
    base = st_idx + text_len; the next chunk starts at max + 1; padded slots hold 1."""
    input_ids = np.asarray(input_ids)
    B, S = input_ids.shape
    if attention_mask is None:
        attention_mask = np.ones_like(input_ids)
    attention_mask = np.asarray(attention_mask)
    if image_grid_thw is None or len(image_grid_thw) == 0:
        pos = np.cumsum(attention_mask.astype(np.int64), axis=-1) - 1
        pos[attention_mask == 0] = 1
        pos3 = np.broadcast_to(pos[None], (3, B, S)).copy()
        return pos3, (pos3.max(axis=(0, 2)) + 1 - S).reshape(B, 1)
    grid = np.asarray(image_grid_thw, dtype=np.int64).reshape(-1, 3)
    pos3 = np.ones((3, B, S), dtype=np.int64)
    deltas = np.zeros((B, 1), dtype=np.int64)
    img = 0
    for b in range(B):
        keep = attention_mask[b] == 1
        ids = input_ids[b][keep]
        L = len(ids)
        starts = np.nonzero(ids == vision_start_token_id)[0]
        starts = starts[starts + 1 < L]
        n_img = int((ids[starts + 1] == image_token_id).sum())
        chunks = []
        st, nxt = 0, 0
        img_positions = np.nonzero(ids == image_token_id)[0]
        for _ in range(n_img):
            cand = img_positions[img_positions >= st]
            ed = int(cand[0])
            t, h, w = (int(v) for v in grid[img])
            img += 1
            gh, gw = h // spatial_merge_size, w // spatial_merge_size
            text_len = ed - st
            chunks.append(np.broadcast_to(np.arange(text_len)[None] + nxt, (3, text_len)))
            base = nxt + text_len
            tt = np.zeros(t * gh * gw, dtype=np.int64)
            hh = np.broadcast_to(np.arange(gh)[None, :, None], (t, gh, gw)).reshape(-1)
            ww = np.broadcast_to(np.arange(gw)[None, None, :], (t, gh, gw)).reshape(-1)
            chunks.append(np.stack([tt, hh, ww]) + base)
            nxt = int(chunks[-1].max()) + 1  # st_idx of the next chunk = max of the last appended chunk + 1 (ref :385)
            st = ed + t * gh * gw
        if st < L:
            chunks.append(np.broadcast_to(np.arange(L - st)[None] + nxt, (3, L - st)))
        p = np.concatenate(chunks, axis=1) if chunks else np.zeros((3, 0), dtype=np.int64)
        pos3[:, b, keep] = p
        deltas[b, 0] = (int(p.max()) + 1 if p.size else 0) - S
    return pos3, deltas


def logprob_rows(input_ids: np.ndarray, response_mask: np.ndarray, flat_index: np.ndarray):
    """Row selection of the fused ``compute_log_probs`` path (Engine.forward_log_probs).  The reference computes
    ``log_softmax(logits[:, :-1])`` everywhere and multiplies by ``response_mask[:, 1:]`` (strategy.py:123-138); here only the
    positions s with ``response_mask[b, s + 1] != 0`` that own a packed hidden row go through the LM head.
    ``flat_index`` [T]: index into the flattened [B * S] grid of each packed (non-padding) token.
    Returns (rows [R] packed hidden-row indices, labels [R] = input_ids[b, s + 1], out_index [R] = b * (S - 1) + s)."""
    B, S = input_ids.shape
    row_of = np.full(B * S, -1, dtype=np.int64)          # padded grid position -> packed row
    row_of[flat_index] = np.arange(len(flat_index), dtype=np.int64)
    row_of = row_of.reshape(B, S)
    want = np.zeros((B, S), dtype=bool)
    want[:, :-1] = np.asarray(response_mask)[:, 1:] != 0   # position s predicts token s + 1
    want &= row_of >= 0                                    # (a masked-in position on padding has no hidden row)
    bs, ss = np.nonzero(want)
    return row_of[bs, ss], np.asarray(input_ids)[bs, ss + 1].astype(np.int64), (bs * (S - 1) + ss).astype(np.int64)
