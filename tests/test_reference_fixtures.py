"""CPU: the oracle's restatements of the reference's in-tree algorithms -- and the product's host logic that mirrors them --
against outputs of the REFERENCE ITSELF (tests/golden/reference_intree.pt, produced by tools/make_reference_fixtures.py, which
executes the reference's own function sources extracted with `ast`; file:line list in that script).  Bit-exact for ids,
indices, masks, byte buffers; 1e-6 for the fp32 log-softmax / entropy reductions."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import hf_oracle as O
from oracle import host_ref
from socioreasoner_b200 import plan
from socioreasoner_b200.roll_shim import collate_fn_to_dict_list
from socioreasoner_b200.roll_shim.send_recv import RecvBucketManager
from socioreasoner_b200.roll_shim.strategy import B200InferStrategy, get_dist_info_from_comm_plan


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN_DIR, "reference_intree.pt"), weights_only=False)


def test_rope_index_restatement_and_product_planner_match_the_in_tree_reference(fx):
    """mcore_adapter/.../modeling_qwen2_5_vl.py:319-441 (images: temporal index = span start; left padding; text only)."""
    assert len(fx["rope_index"]) == 4
    for c in fx["rope_index"]:
        pos, deltas = host_ref.get_rope_index(c["input_ids"], c["image_grid_thw"], c["attention_mask"])
        assert torch.equal(pos, c["position_ids"]) and torch.equal(deltas.view(-1), c["rope_deltas"].view(-1))
        grid = None if c["image_grid_thw"] is None else c["image_grid_thw"].numpy()
        p2, d2 = plan.rope_index(c["input_ids"].numpy(), grid, c["attention_mask"].numpy(), 151655, 151652)
        assert np.array_equal(p2, c["position_ids"].numpy()) and np.array_equal(d2.reshape(-1), c["rope_deltas"].numpy().reshape(-1))


def test_mrope_table_and_apply_restatements_match_the_in_tree_reference(fx):
    """mcore_adapter/.../modeling_qwen2_5_vl.py:14-75: Qwen2_5_VLRotaryEmbedding.forward + apply_multimodal_rotary_pos_emb
    (reference layout [s, b, heads, dim]; the restatement works on [b, s, heads, dim] with the 64-slot half table)."""
    c = fx["mrope"]
    cos, sin = host_ref.mrope_cos_sin(c["position_ids"], c["inv_freq"], c["mrope_section"])
    emb = c["emb"]                                                                  # [3, B, S, 128] = cat(freqs, freqs)
    ref_sel = torch.cat([emb[0, :, :, :16], emb[1, :, :, 16:40], emb[2, :, :, 40:64]], -1)
    assert torch.allclose(cos, ref_sel.cos(), atol=1e-6) and torch.allclose(sin, ref_sel.sin(), atol=1e-6)
    q = c["q"].permute(1, 0, 2, 3)                                                   # [s, b, h, d] -> [b, s, h, d]
    k = c["k"].permute(1, 0, 2, 3)
    assert torch.allclose(host_ref.apply_mrope(q, cos, sin), c["q_embed"].permute(1, 0, 2, 3), atol=1e-5)
    assert torch.allclose(host_ref.apply_mrope(k, cos, sin), c["k_embed"].permute(1, 0, 2, 3), atol=1e-5)


def test_log_prob_and_entropy_restatements_match_the_reference(fx):
    """roll/utils/functionals.py:193-205 and the masking wrappers roll/distributed/strategy/strategy.py:123-138."""
    c = fx["logprob"]
    lg, ids, rm = c["logits"], c["input_ids"], c["response_mask"]
    assert torch.allclose(host_ref.log_probs_from_logits(lg[:, :-1], ids[:, 1:]), c["log_probs_from_logits"], atol=1e-6)
    assert torch.allclose(host_ref.entropy_from_logits(lg), c["entropy_from_logits"], atol=1e-6)
    assert torch.allclose(host_ref.op_compute_log_probs(lg, ids, rm), c["op_compute_log_probs"], atol=1e-6)
    assert torch.allclose(host_ref.op_compute_entropy(lg, rm), c["op_compute_entropy"], atol=1e-6)
    # the fused path's row selection reproduces the reference output from per-row results
    B, S, V = lg.shape
    flat = np.arange(B * S)
    rows, labels, out = plan.logprob_rows(ids.numpy(), rm.numpy(), flat)
    lp = torch.zeros(B * (S - 1))
    lp[out] = host_ref.log_probs_from_logits(lg.reshape(-1, V)[rows], torch.from_numpy(labels))
    assert torch.allclose(lp.view(B, S - 1), c["op_compute_log_probs"], atol=1e-6)


def test_postprocess_generate_restatement_matches_the_reference(fx):
    """roll/utils/functionals.py:768-872, M-RoPE branch (3-axis ids with distinct axes, left-padded prompts, n = 2)."""
    c = fx["postprocess_generate"]
    got = host_ref.postprocess_generate_ref(c["input_ids"], c["attention_mask"], c["position_ids"], c["output"],
                                            c["num_return_sequences"], c["sequence_length"], c["pad_token_id"])
    assert set(got) == set(c["result"])
    for k, v in c["result"].items():
        assert got[k].shape == v.shape and torch.equal(got[k].to(v.dtype), v), k


def test_comm_plan_lookup_and_collate_match_the_reference(fx):
    """roll/utils/functionals.py:875-882; roll/datasets/collator.py:13-40."""
    c = fx["comm_plan"]
    for (rank_in_cluster, rank_in_worker), want in c["lookups"]:
        assert get_dist_info_from_comm_plan(c["plan"], rank_in_cluster, rank_in_worker)[0] == want
    got = collate_fn_to_dict_list(fx["collate"]["inp"])
    want = fx["collate"]["out"]
    assert set(got) == set(want) and torch.equal(got["a"], want["a"])
    assert got["t"].dtype == want["t"].dtype == object and list(got["t"]) == list(want["t"])


def test_giou_restatement_matches_the_reference(fx):
    """roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py:45-58 (0/255-valued masks, both-empty -> 1.0)."""
    c = fx["giou"]
    for (a, b), want in zip(c["pairs"], c["values"]):
        assert host_ref.compute_giou(c["masks"][a], c["masks"][b]) == want


def test_mask_epilogue_restatement_matches_the_reference_segment(fx):
    """roll/distributed/strategy/seg_strategy.py:26-72 executed around a stub predictor: best-score pick, OR over objects,
    objects whose predictor call fails are skipped, cv2 nearest resize 756 -> 768, empty prompt list -> zero 768 mask."""
    for c in fx["seg_epilogue"]:
        want = np.unpackbits(c["mask_bits"])[: 768 * 768].reshape(768, 768).astype(np.uint8)
        if c["tile"] is None:
            got = host_ref.mask_epilogue(torch.zeros(0, 3, 256, 256), torch.zeros(0, 3))
            assert c["shape"] == (768, 768)
        else:
            ml, ms, _ = O.synth_mask_inputs(c["tile"], c["n_obj"])
            got = host_ref.mask_epilogue(ml[c["keep"]], ms[c["keep"]])
        assert got.dtype == np.uint8 and got.shape == (768, 768) and int(got.max()) == c["mask_max"]
        assert np.array_equal(got, want)


def test_bucket_receiver_reassembles_what_the_reference_sender_emits(fx):
    """roll/utils/send_recv_utils.py:64-163: buckets produced by the reference's own SendBucketManager / TensorBucket (tensors
    spanning several 10 kB buckets, dict-form tensor_meta) -> the product's receiver, and through the strategy hook."""
    c = fx["buckets"]
    assert len(c["buckets"]) > 3 and all(b.numel() == c["bucket_size"] and b.dtype == torch.int8 for _, b in c["buckets"])
    rx, got = RecvBucketManager(), {}
    for meta, buf in c["buckets"]:
        got.update(rx.process_bucket(meta, buf))
    rx.clear()
    assert list(got) == [n for n, _ in c["named"]]
    for n, t in c["named"]:
        assert got[n].dtype == t.dtype and torch.equal(got[n], t)
    # the oracle's sender restatement emits byte-identical buckets
    mine = host_ref.send_buckets(c["named"], c["bucket_size"])
    assert len(mine) == len(c["buckets"])
    for (m1, b1), (m2, b2) in zip(mine, c["buckets"]):
        assert {k: (v["bucket_start"], v["tensor_start"], v["save_bytes"]) for k, v in m1.items()} == \
               {k: (v["bucket_start"], v["tensor_start"], v["save_bytes"]) for k, v in m2.items()}
        used = max(v["bucket_start"] + v["save_bytes"] for v in m2.values())
        assert torch.equal(b1[:used], b2[:used])

    class Rec:
        device = "cpu"
        got = {}

        def load_weights(self, named):
            self.got.update({k: v.clone() for k, v in named})

    from types import SimpleNamespace
    st = B200InferStrategy(SimpleNamespace(worker_config=None, rank=0))
    st.model = Rec()
    for meta, buf in c["buckets"]:
        st.update_parameter_in_bucket(meta, buf, ranks_in_worker=[0])
    assert all(torch.equal(st.model.got[n], t) for n, t in c["named"])
