"""Data-parallel rollout over whole tiles + the single result exchange.

Reference behaviour being replaced:
  * ``dispatch_dp_mp_compute`` splits the batch along dim 0 into ``dp_size`` chunks with ``np.array_split`` sizes
    (``roll/distributed/scheduler/decorator.py:106-148``, ``protocol.py:550-617``), one chunk per GPU worker, every worker
    holding a full weight replica; there is no communication inside the forward.
  * ``collect_dp_mp_compute`` + ``DataProto.concat`` bring the per-rank results back through the Ray object store
    (pickled ``torch.save``; ``decorator.py:151-180``, ``protocol.py:718-732``).
Here: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Each rank runs its shard through the engine;
the mask epilogue writes straight into this rank's slot of one registered gather buffer and ONE ``ncclAllGather``
moves masks + IoU + generated tokens of all ranks (rank order == chunk order == the reference's concat order).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import lib


def shard_indices(batch: int, world: int) -> List[np.ndarray]:
    """Chunk sizes/order of ``DataProto.chunk`` (``np.array_split`` semantics, protocol.py:562-566)."""
    return [np.asarray(c, dtype=np.int64) for c in np.array_split(np.arange(batch), world)]


@dataclass
class TileBatch:
    """One rank's shard (host or device tensors)."""
    input_ids: torch.Tensor          # [b,S] int64 (left padded)
    attention_mask: torch.Tensor     # [b,S]
    position_ids: Optional[torch.Tensor]  # [3,b,S] or None (-> reference get_rope_index semantics)
    pixel_values: Optional[torch.Tensor]  # [sum N, 1176] fp32 (HF processor rows), or None when raw tiles are given
    image_grid_thw: torch.Tensor     # [b,3] (cpu)
    mask_logits: torch.Tensor        # [b, n_obj, 3, h, w] fp32  (SAM-style low-res logits)
    mask_scores: torch.Tensor        # [b, n_obj, 3] fp32
    gt_masks: torch.Tensor           # [b, out, out] uint8
    tiles_u8: Optional[torch.Tensor] = None  # [n, H, W, 3] uint8 raw tiles (or a list for mixed sizes): preprocessed ON THE DEVICE when
                                             # pixel_values is None -- 3 MB instead of 13.7 MB of fp32 pixel rows per 1024^2 tile

    def pin(self) -> "TileBatch":
        def f(t):
            if isinstance(t, (list, tuple)):
                return [f(x) for x in t]
            return t.pin_memory() if (t is not None and not t.is_cuda and torch.cuda.is_available()) else t
        return TileBatch(f(self.input_ids), f(self.attention_mask), f(self.position_ids), f(self.pixel_values),
                         self.image_grid_thw, f(self.mask_logits), f(self.mask_scores), f(self.gt_masks), f(self.tiles_u8))

    def h2d_bytes(self) -> int:
        ts = [self.input_ids, self.attention_mask, self.position_ids, self.pixel_values, self.mask_logits, self.mask_scores, self.gt_masks,
              self.tiles_u8]
        flat = []
        for t in ts:
            flat.extend(t if isinstance(t, (list, tuple)) else [t])
        return int(sum(t.numel() * t.element_size() for t in flat if t is not None and not t.is_cuda))


@dataclass
class RolloutResult:
    tokens: torch.Tensor  # [B,T] int64
    masks: torch.Tensor   # [B,out,out] uint8
    iou: torch.Tensor     # [B] fp32


class PendingResult:
    """A result read-back in flight (``DataParallelRollout.gather_to_host_async``).  ``wait()`` blocks until the pinned host
    buffer is complete and returns host (tokens, masks, iou): lists of per-rank views in rank order (no host-side copy;
    ``torch.cat`` them if one tensor is needed), valid until the buffer is reused two read-backs later."""

    def __init__(self, rollout, hbuf, done, b_max):
        self._ro, self._hbuf, self._done, self._b_max = rollout, hbuf, done, b_max

    def wait(self):
        if self._done is not None:
            self._done.synchronize()
        return self._ro._unpack(self._hbuf, self._b_max, cat=False)


class DataParallelRollout:
    def __init__(self, model, max_new_tokens: int = 32, image_size: int = 756, out_size: int = 768, group=None,
                 max_pixels: Optional[int] = None):
        self.model, self.engine = model, (model.engine if model is not None else None)
        self.T, self.image_size, self.out_size = max_new_tokens, image_size, out_size
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._gbuf = None
        self._slot_bytes = 0
        self._pre = None
        self.max_pixels = max_pixels
        # double-buffered upload of the NEXT step's host inputs on a copy stream while the current step computes
        self._copy_stream = None
        self._sets = [None, None]          # device copies (TileBatch) of two host batches
        self._set_done = [None, None]      # event: last device use of that set
        self._pref = None                  # (host batch, set index, copy-finished event)
        # pipelined result read-back (rank 0): gather buffer -> one of two pinned host buffers on a copy stream
        self._d2h_stream = None
        self._d2h_done = None              # event: the last read-back of the gather buffer has finished
        self._hbufs = [None, None]
        self._h_i = 0
        # software pipeline over consecutive batches (submit / flush)
        self._engines = None               # [engine, a second Engine over the same weights]
        self._n_submitted = 0
        self._inflight = None

    # layout of one rank's slot in the gather buffer: [b_max masks][b_max iou f32][b_max*T tokens i64][count i64]
    def _slot_layout(self, b_max: int):
        mask_b = b_max * self.out_size * self.out_size
        iou_off = (mask_b + 15) // 16 * 16
        tok_off = iou_off + (b_max * 4 + 15) // 16 * 16
        cnt_off = tok_off + b_max * self.T * 8
        total = (cnt_off + 8 + 15) // 16 * 16
        return mask_b, iou_off, tok_off, cnt_off, total

    def _ensure_gbuf(self, b_max: int, device):
        *_, total = self._slot_layout(b_max)
        if self._gbuf is None or self._slot_bytes != total:
            self._gbuf = torch.zeros((self.world, total), dtype=torch.uint8, device=device)
            self._slot_bytes = total
        return self._gbuf

    def resolve_b_max(self, b: int, b_max: Optional[int]) -> int:
        """The gather slots of all ranks must have ONE size: with ``np.array_split`` shards the ranks hold different b, so an
        unspecified ``b_max`` is agreed on with an all-reduce(MAX) instead of defaulting to the local shard size."""
        if b_max is not None:
            return int(b_max)
        if self.world == 1:
            return max(b, 1)
        dev = self.engine.device if self.engine is not None else "cpu"
        t = torch.tensor([b], dtype=torch.int64, device=dev if (dist.get_backend(self.group) == "nccl") else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return max(int(t.item()), 1)

    def prefetch(self, batch: TileBatch) -> None:
        """Start copying a HOST batch (pinned) to the device on the copy stream; the next ``step(batch)`` with the same object
        uses the device copy.  Call it right after ``step`` of the current batch: the upload then overlaps the current step's
        compute instead of preceding the next step's first kernel."""
        dev = self.engine.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        i = 0 if (self._pref is None or self._pref[1] == 1) else 1
        cs = self._copy_stream
        if self._set_done[i] is not None:
            cs.wait_event(self._set_done[i])          # the set's previous contents are no longer being read
        with torch.cuda.stream(cs):
            up = lambda t: None if t is None else ([x.to(dev, non_blocking=True) for x in t] if isinstance(t, (list, tuple))
                                                   else (t if t.is_cuda else t.to(dev, non_blocking=True)))
            self._sets[i] = TileBatch(batch.input_ids, batch.attention_mask, batch.position_ids, up(batch.pixel_values), batch.image_grid_thw,
                                      up(batch.mask_logits), up(batch.mask_scores), up(batch.gt_masks), tiles_u8=up(batch.tiles_u8))
            ev = torch.cuda.Event()
            ev.record(cs)
        self._pref = (batch, i, ev)

    def step(self, batch: TileBatch, b_max: Optional[int] = None) -> RolloutResult:
        """Runs this rank's shard.  Host tensors are copied to the device here (pinned -> async H2D); raw uint8 tiles are
        resized / normalised / patchified on the device.  An empty shard (b == 0) only publishes count 0."""
        return self._back(self._front(batch, b_max, self.engine, None, defer=False))

    # ---- software-pipelined form: the decode of batch k is interleaved with ViT + prefill of batch k + 1
    def submit(self, batch: TileBatch, b_max: Optional[int] = None) -> Optional[RolloutResult]:
        """Pipelined ``step``: submits ViT + prefill of ``batch`` with the decode steps of the PREVIOUSLY submitted batch spread
        between its layers, then finishes that previous batch (last decode steps, mask epilogue, IoU, gather slot) and returns its
        result -- None on the first call; ``flush()`` finishes the last batch.  Same results as ``step`` batch for batch.

        Why: ViT + prefill run at the board's power cap and leave the SM clock low; a decode stage that follows as one block pays
        for it for tens of milliseconds ("clock after the prefill").  Alternating the two kinds of work keeps
        the load -- and the clock -- even: 160 GEMM launches + 31 decode steps take 138 ms back to back and 131 ms interleaved
        (tools/decode_after_gemm_probe.py).  Consecutive batches live on two ``Engine`` objects over the SAME weights (buffers,
        KV cache, decode state and captured graphs belong to an engine).  Do not call ``step`` while a batch is in flight."""
        if self._engines is None:
            from .engine import Engine
            e0 = self.engine
            e1 = Engine(e0.cfg, e0.w, e0.device)
            e1.use_persistent_decode, e1.use_cuda_graph, e1.vit_narrow_block_n = e0.use_persistent_decode, e0.use_cuda_graph, e0.vit_narrow_block_n
            self._engines = [e0, e1]
        eng = self._engines[self._n_submitted % 2]
        self._n_submitted += 1
        prev, self._inflight = self._inflight, None
        tick = None
        if prev is not None and prev.get("pending") is not None:
            with_images = batch.pixel_values is not None or batch.tiles_u8 is not None
            n_ticks = eng.prefill_ticks(with_images)
            n_vit = n_ticks - len(eng.w.layers)
            where = os.environ.get("SR_PIPELINE_TICKS", "even")   # experiment knob: "vit" / "dec" = only between ViT blocks / decoder layers
            first, last = (0, n_vit) if (where == "vit" and n_vit > 0) else ((n_vit, n_ticks) if where == "dec" else (0, n_ticks))
            tick = prev["pending"].ticker(n_ticks, first, last)
        self._inflight = self._front(batch, b_max, eng, tick, defer=True)
        return self._back(prev) if prev is not None else None

    def flush(self) -> Optional[RolloutResult]:
        """Finishes the batch still in flight after the last ``submit`` (None if there is none)."""
        prev, self._inflight = self._inflight, None
        return self._back(prev) if prev is not None else None

    def _front(self, batch: TileBatch, b_max: Optional[int], eng, tick, defer: bool) -> dict:
        """Upload (or take the prefetched copy), device preprocessing, ViT + prefill; the decode is deferred when ``tick`` / a
        pipelined engine asks for it."""
        dev = self.engine.device
        used_set = None
        if self._pref is not None and self._pref[0] is batch:       # inputs were uploaded ahead of time
            _, used_set, ev = self._pref
            torch.cuda.current_stream(dev).wait_event(ev)
            batch = self._sets[used_set]
        b = batch.input_ids.shape[0]
        b_max = self.resolve_b_max(b, b_max)
        assert b <= b_max
        self._ensure_gbuf(b_max, dev)
        fl = dict(b=b, b_max=b_max, used_set=used_set, pending=None, tokens=None)
        if b == 0:
            return fl
        to = lambda t: t if (t is None or t.is_cuda) else t.to(dev, non_blocking=True)
        pv, ml, ms, gt = to(batch.pixel_values), to(batch.mask_logits), to(batch.mask_scores), to(batch.gt_masks)
        grid = batch.image_grid_thw
        if pv is None:
            if self._pre is None:
                from .preprocess import GpuPreprocessor, REFERENCE_MAX_PIXELS
                self._pre = GpuPreprocessor(dev, max_pixels=self.max_pixels or REFERENCE_MAX_PIXELS)
            tiles = [to(t) for t in batch.tiles_u8] if isinstance(batch.tiles_u8, (list, tuple)) else to(batch.tiles_u8)
            pv, grid2 = self._pre([tiles[i] for i in range(len(tiles))], out_dtype=torch.bfloat16)
            assert grid is None or grid2.tolist() == [list(map(int, g)) for g in grid.tolist()], "tile sizes do not match image_grid_thw"
            grid = grid2
        pipelined = defer and self.T > 1 and eng.use_cuda_graph and lib.timeline is None
        out = eng.generate_greedy(batch.input_ids, batch.attention_mask, batch.position_ids, pv, grid, self.T, tick=tick,
                                  defer_decode=pipelined)
        fl.update(ml=ml, ms=ms, gt=gt)
        if pipelined:
            fl["pending"] = out
        else:
            fl["tokens"] = out
        return fl

    def _back(self, fl: dict) -> RolloutResult:
        """Remaining decode steps, mask epilogue straight into this rank's gather slot, IoU, slot packing."""
        dev = self.engine.device
        b, b_max = fl["b"], fl["b_max"]
        if b == 0:
            self.pack_slot(b_max, torch.zeros((0, self.T), dtype=torch.int64, device=dev), torch.zeros((0,), dtype=torch.float32, device=dev))
            return RolloutResult(torch.zeros((0, self.T), dtype=torch.int64, device=dev), self.slot_masks(b_max)[:0],
                                 torch.zeros((0,), dtype=torch.float32, device=dev))
        tokens = fl["pending"].finish() if fl["pending"] is not None else fl["tokens"]
        masks = self.slot_masks(b_max)
        if self._d2h_done is not None:     # the previous step's read-back of the gather buffer must finish before it is rewritten
            torch.cuda.current_stream(dev).wait_event(self._d2h_done)
            self._d2h_done = None
        ml, ms, gt = fl["ml"], fl["ms"], fl["gt"]
        for i in range(b):
            lib.mask_epilogue(ml[i], ms[i], self.image_size, self.out_size, out=masks[i])
        iou = lib.mask_iou(masks[:b], gt, 1.0)
        self.pack_slot(b_max, tokens, iou)
        if fl["used_set"] is not None:
            self._set_done[fl["used_set"]] = torch.cuda.Event()
            self._set_done[fl["used_set"]].record(torch.cuda.current_stream(dev))
        return RolloutResult(tokens, masks[:b], iou)

    def slot_masks(self, b_max: int, device=None) -> torch.Tensor:
        """This rank's mask region of the gather buffer ([b_max, out, out] uint8 view): the epilogue's output target."""
        gbuf = self._ensure_gbuf(b_max, device if device is not None else self.engine.device)
        mask_b = self._slot_layout(b_max)[0]
        return gbuf[self.rank][:mask_b].view(b_max, self.out_size, self.out_size)

    def pack_slot(self, b_max: int, tokens: torch.Tensor, iou: torch.Tensor) -> None:
        """Write IoU, tokens and the valid-count next to the masks in this rank's slot."""
        b = tokens.shape[0]
        _, iou_off, tok_off, cnt_off, _ = self._slot_layout(b_max)
        slot = self._gbuf[self.rank]
        slot[iou_off:iou_off + b * 4].view(torch.float32).copy_(iou)
        slot[tok_off:tok_off + b * self.T * 8].view(torch.int64).copy_(tokens.reshape(-1))
        slot[cnt_off:cnt_off + 8].view(torch.int64).fill_(b)

    def gather_to_host(self, b_max: int, dst_rank: int = 0):
        """``gather`` + ONE device -> host copy of the whole gather buffer into a PINNED host buffer on ``dst_rank`` only (the
        reference collects on the driver only, decorator.py:151-180); the other ranks return None.  Returns host
        (tokens, masks, iou): per-rank VIEWS of that buffer concatenated lazily -- lists of tensors in rank order, no host-side copy
        (``torch.cat`` them if one tensor is needed); valid until the next call."""
        pend = self.gather_to_host_async(b_max, dst_rank)
        return None if pend is None else pend.wait()

    def gather_to_host_async(self, b_max: int, dst_rank: int = 0) -> Optional["PendingResult"]:
        """``gather_to_host`` without blocking the host: the copy into one of two pinned host buffers runs on a copy stream behind
        the all-gather, so the driver-side consumer reads step k's result (``PendingResult.wait()``) while step k + 1 is already
        running; the next step's epilogue waits for the copy before it rewrites the gather buffer.  None on the other ranks."""
        self._all_gather()
        if self.rank != dst_rank:
            return None
        gbuf = self._gbuf
        self._h_i ^= 1
        hb = self._hbufs[self._h_i]
        if hb is None or hb.shape != gbuf.shape:
            hb = self._hbufs[self._h_i] = torch.empty(gbuf.shape, dtype=torch.uint8, pin_memory=gbuf.is_cuda)
        if not gbuf.is_cuda:
            hb.copy_(gbuf)
            return PendingResult(self, hb, None, b_max)
        dev = gbuf.device
        if self._d2h_stream is None:
            self._d2h_stream = torch.cuda.Stream(device=dev)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        self._d2h_stream.wait_event(ready)
        with torch.cuda.stream(self._d2h_stream):
            hb.copy_(gbuf, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._d2h_stream)
        self._d2h_done = done
        return PendingResult(self, hb, done, b_max)

    def gather(self, b_max: int) -> RolloutResult:
        """ONE all-gather of every rank's slot; returns the concatenation in rank order (= reference concat order)."""
        self._all_gather()
        return self._unpack(self._gbuf, b_max)

    def _all_gather(self) -> None:
        gbuf = self._gbuf
        if self.world > 1:
            if gbuf.is_cuda:  # NCCL in-place all-gather (send buffer = this rank's slot of the receive buffer)
                dist.all_gather_into_tensor(gbuf.view(-1), gbuf[self.rank], group=self.group)
            else:             # gloo (CPU tests of the host logic)
                parts = [torch.empty_like(gbuf[0]) for _ in range(self.world)]
                dist.all_gather(parts, gbuf[self.rank].clone(), group=self.group)
                gbuf.copy_(torch.stack(parts))

    def _unpack(self, gbuf: torch.Tensor, b_max: int, cat: bool = True):
        mask_b, iou_off, tok_off, cnt_off, _ = self._slot_layout(b_max)
        counts = gbuf[:, cnt_off:cnt_off + 8].contiguous().view(torch.int64).view(-1).tolist()
        toks, masks, ious = [], [], []
        for r, c in enumerate(counts):
            c = int(c)
            masks.append(gbuf[r, :mask_b].view(b_max, self.out_size, self.out_size)[:c])
            ious.append(gbuf[r, iou_off:iou_off + b_max * 4].view(torch.float32)[:c])
            toks.append(gbuf[r, tok_off:tok_off + b_max * self.T * 8].view(torch.int64).view(b_max, self.T)[:c])
        if not cat:
            return toks, masks, ious
        return RolloutResult(torch.cat(toks), torch.cat(masks), torch.cat(ious))
