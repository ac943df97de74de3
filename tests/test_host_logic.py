"""CPU: host-side logic of the product package against the oracle / HF reference implementations."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import hf_oracle as O
from oracle import host_ref
from socioreasoner_b200 import plan, preprocess, weights
from socioreasoner_b200.config import SocioReasonerConfig
from socioreasoner_b200.engine import Engine
from socioreasoner_b200.rollout import shard_indices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("grid", [[[1, 32, 32]], [[1, 54, 54]], [[1, 16, 16], [1, 32, 64], [1, 74, 74]], [[1, 96, 96]],
                                  [[1, 2, 2]], [[1, 8, 4], [1, 4, 8]], [[2, 16, 16]]])
def test_vit_plan_matches_hf(oracle_tiny, grid):
    vis = oracle_tiny.model.visual
    g = torch.tensor(grid)
    wi, cu = vis.get_window_index(g)
    cu = torch.unique_consecutive(torch.tensor(cu))
    vp = plan.vit_plan(np.array(grid))
    assert np.array_equal(vp.cu_window_seqlens, cu.numpy())
    perm = (wi[:, None] * 4 + torch.arange(4)[None]).reshape(-1).numpy()
    assert np.array_equal(vp.row_perm, perm)
    assert np.array_equal(vp.merged_reverse, torch.argsort(wi).numpy())
    # rotary position ids: HF rot_pos_emb builds freqs[pos_ids]; compare through the frequency table
    rpe = vis.rot_pos_emb(g).reshape(-1, 4, 40)[wi].reshape(-1, 40)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 40, 2, dtype=torch.float) / 40))
    mine = torch.cat([torch.from_numpy(vp.hpos).float()[:, None] * inv[None], torch.from_numpy(vp.wpos).float()[:, None] * inv[None]], 1)
    assert torch.equal(rpe, mine)
    # attention plans: every row's key range == its window / image, tiles cover all rows exactly once
    for p, bounds in ((vp.window_plan, vp.cu_window_seqlens), (vp.full_plan, vp.cu_seqlens)):
        seg = np.searchsorted(bounds, np.arange(vp.n), side="right") - 1
        assert np.array_equal(p.row_lo, bounds[seg]) and np.array_equal(p.row_hi, bounds[seg + 1])
        cover = np.zeros(vp.n, int)
        for q0, nr, lo, hi in p.tiles:
            assert 0 < nr <= 128 and lo <= p.row_lo[q0:q0 + nr].min() and hi >= p.row_hi[q0:q0 + nr].max()
            cover[q0:q0 + nr] += 1
        assert np.all(cover == 1)


def test_window_plan_packs_whole_windows():
    """ViT window tiles are runs of whole windows: every tile's key range is exactly its own rows (one key block), rows are covered
    once, and a segment longer than a tile is chunked with the whole segment as key range."""
    vp = plan.vit_plan(np.array([[1, 54, 54]] * 2 + [[1, 32, 32]]))
    wp = vp.window_plan
    cov = np.zeros(vp.n, dtype=int)
    bounds = set(vp.cu_window_seqlens.tolist())
    for r0, nr, lo, hi in wp.tiles.tolist():
        cov[r0:r0 + nr] += 1
        assert nr <= 128 and (lo, hi) == (r0, r0 + nr) and r0 in bounds and r0 + nr in bounds
    assert (cov == 1).all()
    blind = plan.plan_from_segments(vp.cu_window_seqlens, vp.n)          # round-1 tiling: 128-row chunks regardless of windows
    blocks = lambda t: int(((t[:, 3] - t[:, 2] + 127) // 128).sum())
    assert blocks(wp.tiles) < 0.7 * blocks(blind.tiles)
    assert np.array_equal(wp.row_lo, blind.row_lo) and np.array_equal(wp.row_hi, blind.row_hi)
    long = plan.plan_from_segments(np.array([0, 40, 340, 400]), 400, pack_whole=True)
    assert long.tiles.tolist() == [[0, 40, 0, 40], [40, 128, 40, 340], [168, 128, 40, 340], [296, 44, 40, 340], [340, 60, 340, 400]]


def test_causal_plan():
    lens, L = [5, 300, 128], 384
    qs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ks = np.arange(3) * L
    p = plan.plan_causal(lens, qs, ks)
    assert len(p.row_lo) == sum(lens)
    for b, (n, q0, k0) in enumerate(zip(lens, qs, ks)):
        assert np.all(p.row_lo[q0:q0 + n] == k0) and np.array_equal(p.row_hi[q0:q0 + n], k0 + 1 + np.arange(n))
    want = [(0, 5, 0, 5), (5, 128, 384, 512), (133, 128, 384, 640), (261, 44, 384, 684), (305, 128, 768, 896)]
    assert sorted(tuple(int(v) for v in t) for t in p.tiles) == want
    spans = [int(t[3] - t[2]) for t in p.tiles]
    assert spans == sorted(spans, reverse=True)   # longest key range first (scheduling order)


def test_rope_index_matches_oracle():
    for tiles, sides in (([0, 1, 2], [224, 336, 448]), ([3], [224])):
        b = O.make_batch(tiles, sides, "tiny", n_text=16)
        p3, d = plan.rope_index(b.input_ids.numpy(), b.image_grid_thw.numpy(), b.attention_mask.numpy(), 8001, 8003)
        ref, dref = host_ref.get_rope_index(b.input_ids, b.image_grid_thw, b.attention_mask, image_token_id=8001, vision_start_token_id=8003)
        assert np.array_equal(p3, ref.numpy()) and np.array_equal(d, dref.numpy())
    ids = torch.randint(0, 100, (2, 9)); mask = torch.tensor([[0, 0, 0, 1, 1, 1, 1, 1, 1], [1] * 9])
    p3, d = plan.rope_index(ids.numpy(), None, mask.numpy(), 8001, 8003)
    ref, dref = host_ref.get_rope_index(ids, None, mask)
    assert np.array_equal(p3, ref.numpy()) and np.array_equal(d, dref.numpy())


def test_two_images_per_prompt_rope_index():
    # the reference sends map + sat images in one prompt (rlvr_socioseg_vlm_pipeline_infer.py:74-81)
    tok = O.CONFIGS["tiny"]["tokens"]
    ids = torch.tensor([[7001, tok["vision_start_token_id"]] + [tok["image_token_id"]] * 4 + [tok["vision_end_token_id"], 11,
                         tok["vision_start_token_id"]] + [tok["image_token_id"]] * 6 + [tok["vision_end_token_id"], 12, 13]])
    grid = torch.tensor([[1, 4, 4], [1, 4, 6]])
    p3, _ = plan.rope_index(ids.numpy(), grid.numpy(), None, tok["image_token_id"], tok["vision_start_token_id"])
    ref, _ = host_ref.get_rope_index(ids, grid, None, image_token_id=tok["image_token_id"], vision_start_token_id=tok["vision_start_token_id"])
    assert np.array_equal(p3, ref.numpy())


def test_smart_resize_matches_hf():
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize as hf_sr
    for h, w in ((1024, 1024), (448, 448), (1344, 1344), (100, 3000), (756, 1000), (29, 57)):
        for mp in (768 * 768, 1344 * 1344):
            assert preprocess.smart_resize(h, w, 28, 56 * 56, mp) == hf_sr(h, w, factor=28, min_pixels=56 * 56, max_pixels=mp)
    assert preprocess.smart_resize(1024, 1024) == (756, 756)   # the reference-effective tile size


def test_preprocess_equals_oracle_pipeline():
    t = [preprocess.synth_tile(5, 224)]
    pv, grid = preprocess.preprocess_tiles(t)
    pv_o, grid_o = O.preprocess_tiles([O.synth_tile_uint8(5, 224)])
    assert torch.equal(pv, pv_o[0]) and torch.equal(grid, grid_o[0])
    assert torch.equal(preprocess.synth_prompt(3, 16), O.synth_prompt(3, 16, "3b"))


def test_gate_up_interleave_and_padding():
    g = torch.arange(40 * 3, dtype=torch.float32).view(40, 3)
    u = -g
    p = weights.interleave_gate_up(g, u, 48)
    assert p.shape == (96, 3)
    for blk in range(3):
        for j in range(16):
            r = blk * 16 + j
            exp_g = g[r] if r < 40 else torch.zeros(3)
            assert torch.equal(p[blk * 32 + j], exp_g) and torch.equal(p[blk * 32 + 16 + j], -exp_g)


def test_config_accepts_4x_and_5x_layouts():
    c5 = SocioReasonerConfig.from_hf(O.make_config("3b"))
    flat = dict(hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16, num_key_value_heads=2,
                vocab_size=151936, rms_norm_eps=1e-6, rope_theta=1000000.0, tie_word_embeddings=True,
                rope_scaling=dict(type="mrope", mrope_section=[16, 24, 24]),
                vision_config=dict(depth=32, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=2048,
                                   fullatt_block_indexes=[7, 15, 23, 31]),
                image_token_id=151655, vision_start_token_id=151652)
    c4 = SocioReasonerConfig.from_dict(flat)
    assert c4.text == c5.text and c4.vision == c5.vision
    c4.validate()
    assert c4.vision.intermediate_padded == 3456 and c4.vision.patch_dim == 1176


def test_seq_plan_and_shards():
    m = np.array([[0, 0, 1, 1, 1], [1, 1, 1, 1, 1], [0, 0, 0, 0, 1]])
    sp = Engine.seq_plan(m)
    assert sp.lengths.tolist() == [3, 5, 1] and sp.q_starts.tolist() == [0, 3, 8] and sp.total == 9
    assert sp.flat_index.tolist() == [2, 3, 4, 5, 6, 7, 8, 9, 14]
    for batch, world in ((64, 8), (250, 4), (7, 4), (3, 8)):
        assert [c.tolist() for c in shard_indices(batch, world)] == [c.tolist() for c in host_ref.dp_chunk_indices(batch, world)]


def test_logprob_rows_reproduce_the_masked_reference_outputs():
    """Engine.forward_log_probs pushes only the rows selected by plan.logprob_rows through the LM head; scattering their
    per-row results must equal op_compute_log_probs / op_compute_entropy (strategy.py:123-138) on the full logits -- for left
    padding, masks that reach into the padding, an all-zero mask and the last position (which predicts nothing)."""
    g = torch.Generator().manual_seed(0)
    B, S, V = 4, 9, 23
    attn = torch.ones(B, S, dtype=torch.long)
    attn[0, :3] = 0; attn[2, :7] = 0                      # left padding, as the reference collator produces
    ids = torch.randint(0, V, (B, S), generator=g)
    logits_full = torch.randn(B, S, V, generator=g)
    flat = np.nonzero(attn.numpy().reshape(-1))[0]        # packed tokens (SeqPlan.flat_index)
    hidden_logits = logits_full.reshape(-1, V)[flat]      # what the LM head would produce for each packed row
    for rm in (attn.clone(), torch.zeros_like(attn), torch.ones_like(attn),
               torch.tensor([[0, 0, 0, 0, 1, 1, 1, 0, 0]] * B)):
        rows, labels, out = plan.logprob_rows(ids.numpy(), rm.numpy(), flat)
        lp = torch.zeros(B * (S - 1)); ent = torch.zeros(B * (S - 1))
        sel = hidden_logits[rows]
        lp[out] = host_ref.log_probs_from_logits(sel, torch.from_numpy(labels)) if len(rows) else torch.zeros(0)
        ent[out] = host_ref.entropy_from_logits(sel) if len(rows) else torch.zeros(0)
        ref_lp = host_ref.op_compute_log_probs(logits_full, ids, rm)
        ref_ent = host_ref.op_compute_entropy(logits_full, rm)
        # positions that are masked in but sit on padding have no hidden row here; the reference would score its zero-padded
        # logits there -- such masks do not occur (response_mask is a subset of attention_mask), so compare on real tokens
        real = (attn[:, :-1] != 0)
        assert torch.allclose(lp.view(B, S - 1)[real], ref_lp[real], atol=1e-6)
        assert torch.allclose(ent.view(B, S - 1)[real], ref_ent[real], atol=1e-6)
        assert (lp.view(B, S - 1)[~real] == 0).all()
        assert len(set(out.tolist())) == len(out) and (rows >= 0).all() and len(rows) == int(((rm[:, 1:] != 0) & real).sum())


def test_library_exports_every_declared_symbol():
    """The C-ABI shared library loads (no GPU needed) and exports every function include/*.h declares."""
    from socioreasoner_b200 import lib

    hdr = open(os.path.join(ROOT, "include", "socioreasoner_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    if not os.path.exists(lib.LIB_PATH):
        pytest.fail("libsocioreasoner_b200.so missing: run __graft_entry__.build()")
    dll = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.SYMBOLS), (declared ^ set(lib.SYMBOLS))
    assert lib.load().sr_abi_version() == 1


def test_product_path_does_not_import_oracle():
    """No module of the product package (nor bench's own arm) may import oracle/ (it is the checker only)."""
    pkg = os.path.join(ROOT, "socioreasoner_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"


def test_ops_fail_loudly_without_gpu():
    from socioreasoner_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        lib.rmsnorm(torch.zeros(8, 1280, dtype=torch.bfloat16), torch.ones(1280))


def test_pending_decode_ticker_spreads_the_steps_evenly():
    """``PendingDecode.ticker(n)``: the decode steps still to run are submitted evenly over n calls (the pipelined rollout calls it once
    per ViT block / decoder layer of the NEXT batch) and are all done after the n-th call, whatever the ratio."""
    import types
    from socioreasoner_b200.engine import PendingDecode

    class FakeGraph:
        def __init__(self):
            self.replays = 0

        def replay(self):
            self.replays += 1

    for T, n_ticks in ((32, 68), (32, 36), (5, 68), (40, 7), (2, 1)):
        g = FakeGraph()
        eng = types.SimpleNamespace(_decode_launches=3)
        st = {"track_eos": False}
        pd = PendingDecode(eng, st, None, None, g, T, None, 0, 16, None)
        tick = pd.ticker(n_ticks)
        seen = []
        for _ in range(n_ticks):
            tick()
            seen.append(g.replays)
        assert seen[-1] == T - 1 and pd.remaining == 0
        ideal = [(T - 1) * (i + 1) / n_ticks for i in range(n_ticks)]
        assert all(0 <= s_ - x < 1 + 1e-9 for s_, x in zip(seen, ideal))          # never behind, never more than one step ahead
        tick()                                                                     # extra calls are harmless
        assert g.replays == T - 1
