"""GPU (-m gpu): the persistent decode-step kernel (sr_decode_step_persistent) against the per-layer decode kernels --
phase by phase (qkv, attention + KV append, o-proj, SwiGLU, down), a whole step, CUDA-graph replay, ragged KV lengths
(1 .. 1000 cached tokens), batch < 8 -- and end to end through generate_greedy (tokens and logits of both decode paths)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_phases_full_step_and_graph_replay():
    import persistent_bringup as PB

    ok, S, layers, P, ref = PB.check_case(8, 2, [900, 897, 640, 512, 129, 128, 9, 1], 8192)
    assert ok
    ok, _ = PB.graph_check(P, ref)
    assert ok


def test_small_batch_many_splits():
    import persistent_bringup as PB

    ok, *_ = PB.check_case(3, 3, [300, 17, 1000], 4096, phases=False)
    assert ok
    ok, *_ = PB.check_case(1, 1, [77], 4096, phases=False)
    assert ok


def test_generate_greedy_same_tokens_on_both_decode_paths(oracle_tiny, tiny_batch):
    from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration as Model

    model = Model.from_hf_model(oracle_tiny, device="cuda")
    b, T = tiny_batch, 6
    eng = model.engine
    args = (b.input_ids, b.attention_mask, b.position_ids, b.pixel_values, b.image_grid_thw, T)
    eng.set_persistent_decode(False)
    toks0, lg0 = eng.generate_greedy(*args, return_logits=True)
    eng.set_persistent_decode(True)
    assert eng._persistent_ok(b.input_ids.shape[0])
    toks1, lg1 = eng.generate_greedy(*args, return_logits=True)
    toks2 = eng.generate_greedy(*args)  # CUDA-graph replay of the persistent step
    top2 = lg0.topk(2, -1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 0.05
    for i in range(toks0.shape[0]):
        for t in range(T):
            if not decisive[i, t]:
                break
            assert toks0[i, t] == toks1[i, t] == toks2[i, t], (i, t)
    first = lg0[:, 0].float()  # same prefill, same first decode input: logits of step 0 come from the prefill path in both
    assert torch.equal(first, lg1[:, 0].float())
    err = ((lg1[:, 1] - lg0[:, 1]).float().pow(2).mean().sqrt() / lg0[:, 1].float().pow(2).mean().sqrt()).item()
    assert err < 1e-2, err
