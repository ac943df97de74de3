"""CPU: pin the oracle (HF eager + in-tree restatements) against the committed golden vectors."""
import numpy as np
import pytest
import torch

from oracle import hf_oracle as O
from oracle import host_ref


def test_inputs_pinned(golden_tiny, tiny_batch):
    b = tiny_batch
    assert np.array_equal(golden_tiny["grid"], b.image_grid_thw.numpy())
    assert int(golden_tiny["in_ids_sum"]) == int(b.input_ids.sum())
    assert int(golden_tiny["in_pos_sum"]) == int(b.position_ids.sum())
    assert abs(float(golden_tiny["in_pixel_sum"]) - b.pixel_values.double().sum().item()) < 1e-3
    assert abs(float(golden_tiny["in_pixel_abs_sum"]) - b.pixel_values.double().abs().sum().item()) < 1e-3


def test_oracle_matches_golden(golden_tiny, oracle_tiny, tiny_batch):
    g, hf, b = golden_tiny, oracle_tiny, tiny_batch
    merged, hid, caps = O.vit_forward(hf, b.pixel_values, b.image_grid_thw, capture_blocks=True)
    np.testing.assert_allclose(merged[::7, :64].numpy(), g["vit_merged_sub"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(hid[::37, :32].numpy(), g["vit_last_hidden_sub"], rtol=1e-4, atol=1e-4)
    logits = O.forward_logits(hf, b).logits.float()
    keep = b.attention_mask.bool()
    np.testing.assert_allclose(logits[keep][:, ::97].numpy(), g["logits_sub"], rtol=1e-4, atol=2e-4)
    assert np.array_equal(logits[keep].topk(5, -1).indices.numpy()[:, 0], g["logits_top5_idx"][:, 0])
    lp = host_ref.op_compute_log_probs(logits, b.input_ids, b.attention_mask)
    np.testing.assert_allclose(lp.numpy(), g["log_probs"], rtol=1e-4, atol=2e-4)
    ent = host_ref.op_compute_entropy(logits, b.attention_mask)
    np.testing.assert_allclose(ent.numpy(), g["entropy"], rtol=1e-4, atol=2e-4)


def test_oracle_greedy_matches_golden(golden_tiny, oracle_tiny, tiny_batch):
    toks, lg = O.greedy_decode(oracle_tiny, tiny_batch, golden_tiny["greedy_tokens"].shape[1])
    assert np.array_equal(toks.numpy(), golden_tiny["greedy_tokens"])
    np.testing.assert_allclose(lg[:, :, ::97].numpy(), golden_tiny["greedy_logits_sub"], rtol=1e-4, atol=2e-4)


def test_greedy_step0_is_argmax_of_full_forward(oracle_tiny, tiny_batch):
    logits = O.forward_logits(oracle_tiny, tiny_batch).logits
    toks, _ = O.greedy_decode(oracle_tiny, tiny_batch, 1)
    assert torch.equal(toks[:, 0], logits[:, -1].argmax(-1))


def test_rope_index_reference_semantics():
    # image temporal index is 0 => t == base for every image token (reference, not transformers-5.x, semantics)
    ids = torch.tensor([[7001, 8003] + [8001] * 6 + [8004, 5, 6, 7]])
    grid = torch.tensor([[1, 4, 6]])  # merged 2 x 3
    pos, delta = host_ref.get_rope_index(ids, grid, None, image_token_id=8001, vision_start_token_id=8003)
    assert pos[:, 0, :2].tolist() == [[0, 1]] * 3
    assert pos[0, 0, 2:8].tolist() == [2] * 6                      # t
    assert pos[1, 0, 2:8].tolist() == [2, 2, 2, 3, 3, 3]           # h
    assert pos[2, 0, 2:8].tolist() == [2, 3, 4, 2, 3, 4]           # w
    assert pos[:, 0, 8:].tolist() == [[5, 6, 7, 8]] * 3            # next text starts at max + 1
    assert int(delta[0, 0]) == 9 - 12


def test_rope_index_left_padding_and_text_only():
    ids = torch.tensor([[0, 0, 5, 6, 7]])
    mask = torch.tensor([[0, 0, 1, 1, 1]])
    pos, _ = host_ref.get_rope_index(ids, None, mask)
    assert pos[0, 0].tolist() == [1, 1, 0, 1, 2]


def test_compute_giou_cases():
    z = np.zeros((4, 4), np.uint8)
    a = z.copy(); a[:2] = 1
    b = z.copy(); b[1:3] = 255
    assert host_ref.compute_giou(z, z) == 1.0          # both empty -> 1.0 (infer pipeline)
    assert host_ref.compute_iou_reward(z, z) == 0.0     # reward worker variant
    assert host_ref.compute_giou(a, z) == 0.0
    assert abs(host_ref.compute_giou(a, b) - 4 / 12) < 1e-12
    assert host_ref.compute_iou_reward(a, np.zeros((3, 3), np.uint8)) == 0.0


def test_mask_epilogue_oracle_properties():
    ml, ms, gt = O.synth_mask_inputs(0, 3)
    m = host_ref.mask_epilogue(ml, ms)
    assert m.shape == (768, 768) and m.dtype == np.uint8 and set(np.unique(m)) <= {0, 1}
    # OR is monotone in the object set
    m1 = host_ref.mask_epilogue(ml[:1], ms[:1])
    assert np.all(m >= m1)
    # empty prompt list -> zero mask without touching the model (seg_strategy.py:40-45)
    assert host_ref.mask_epilogue(ml[:0], ms[:0]).sum() == 0
    # argmax-score pick: making candidate 2 the best for object 0 changes the result to that candidate
    ms2 = ms.clone(); ms2[0] = torch.tensor([0.0, 0.0, 1.0])
    up = torch.nn.functional.interpolate(ml[0:1], (756, 756), mode="bilinear", align_corners=False)[0, 2] > 0
    import cv2
    exp = cv2.resize(up.numpy().astype(np.uint8), (768, 768), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(host_ref.mask_epilogue(ml[:1], ms2[:1]), exp)


def test_dp_chunking_follows_array_split():
    for batch, world in ((64, 8), (250, 4), (7, 4), (3, 8)):
        chunks = host_ref.dp_chunk_indices(batch, world)
        assert [len(c) for c in chunks] == [len(c) for c in np.array_split(np.arange(batch), world)]
        assert np.array_equal(np.concatenate(chunks), np.arange(batch))
