"""GPU (-m gpu): the reference-facing seam end to end on the real model -- checkpoint directory -> ``create_strategy(worker)``
-> ``initialize(default_actor_model_provider)`` (tokenizer + processor + ``from_pretrained``) -> ``generate`` on the vLLM input
form the unchanged entry script sends (PIL images + prompt ids, rlvr_socioseg_vlm_pipeline_infer.py:645-657) and
``forward_step`` on the HF form -- against the HF oracle loaded from the SAME checkpoint files."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from PIL import Image

import ckpt_util
from oracle import hf_oracle as O
from oracle import host_ref
from socioreasoner_b200.roll_shim import DataProto, create_strategy
from socioreasoner_b200.roll_shim.model_providers import default_actor_model_provider

pytestmark = pytest.mark.gpu
TILES, SIDES = [0, 1, 2], [224, 336, 224]


@pytest.fixture(scope="module")
def setup(tmp_path_factory, oracle_tiny):
    from transformers import Qwen2_5_VLForConditionalGeneration
    path = ckpt_util.write_checkpoint(str(tmp_path_factory.mktemp("ckpt")), oracle_tiny)
    hf = Qwen2_5_VLForConditionalGeneration.from_pretrained(path, attn_implementation="eager", dtype=torch.float32).eval()
    w = SimpleNamespace(worker_config=SimpleNamespace(strategy_args=SimpleNamespace(strategy_name="vllm"),
                                                     model_args=SimpleNamespace(model_name_or_path=path)),
                        rank_info=SimpleNamespace(dp_rank=None, dp_size=None), rank=0)
    st = create_strategy(w)
    st.initialize(default_actor_model_provider)
    b = O.make_batch(TILES, SIDES, "tiny", n_text=[8, 5, 11])       # ragged, left-padded; sides need no resize
    return st, hf, b


def _vllm_form(b, micro=2):
    imgs = [Image.fromarray(np.ascontiguousarray(O.synth_tile_uint8(i, s).transpose(1, 2, 0))) for i, s in zip(TILES, SIDES)]
    mm = [{"prompt_token_ids": b.input_ids[i][b.attention_mask[i].bool()].tolist(), "multi_modal_data": {"image": [imgs[i]]}}
          for i in range(len(imgs))]
    return DataProto.from_dict(tensors=dict(input_ids=b.input_ids, attention_mask=b.attention_mask,
                                            position_ids=b.position_ids.transpose(0, 1).contiguous()),      # stored [B, 3, S]
                               non_tensors=dict(multi_modal_data=mm), meta_info=dict(micro_batch_size=micro))


def test_initialize_built_everything(setup):
    st, hf, b = setup
    assert st.tokenizer is not None and st.tokenizer.eos_token_id == 7002 and st.processor is not None
    assert type(st.model).__name__ == "SocioReasonerB200ForConditionalGeneration"
    assert st.worker.rank_info.dp_size == 1


def test_generate_vllm_form_matches_oracle_greedy(setup):
    st, hf, b = setup
    T = 6
    toks_ref, lg_ref = O.greedy_decode(hf, b, T)
    gc = dict(max_new_tokens=T, temperature=0.0, top_p=1.0, top_k=-1, num_beams=1, repetition_penalty=1.0, num_return_sequences=1,
              eos_token_id=[st.tokenizer.eos_token_id], pad_token_id=st.tokenizer.pad_token_id)
    out = st.generate(_vllm_form(b), gc)
    S = b.input_ids.shape[1]
    assert out.shape[0] == 3 and torch.equal(out[:, :S].cpu(), b.input_ids)
    new = out[:, S:].cpu()
    top2 = lg_ref.topk(2, -1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 0.05   # as tests/test_gpu_e2e.py: below that bf16 cannot decide
    checked = 0
    for i in range(3):
        for t in range(min(T, new.shape[1])):
            if not decisive[i, t] or int(toks_ref[i, t]) == st.tokenizer.eos_token_id:
                break
            assert int(new[i, t]) == int(toks_ref[i, t]), (i, t)
            checked += 1
    assert checked >= 8
    # stop ids: make the oracle's second token of sample 0 a stop id -> its response ends there, padded afterwards
    stop = int(toks_ref[0, 1])
    if decisive[0, 0] and decisive[0, 1]:
        gc2 = dict(gc, eos_token_id=[st.tokenizer.eos_token_id, stop])
        out2 = st.generate(_vllm_form(b), gc2)[:, S:].cpu()
        k = int((toks_ref[0] == stop).nonzero()[0])
        assert int(out2[0, k]) == stop and (out2[0, k + 1:] == st.tokenizer.pad_token_id).all()


def test_forward_step_log_probs_both_paths(setup):
    st, hf, b = setup
    ref_logits = O.forward_logits(hf, b).logits.float()
    rm = b.attention_mask.clone()
    rm[:, : b.input_ids.shape[1] // 2] = 0
    lp_ref = host_ref.op_compute_log_probs(ref_logits, b.input_ids, rm)
    ent_ref = host_ref.op_compute_entropy(ref_logits, rm)
    per = (b.image_grid_thw.prod(-1)).tolist()
    pvs = torch.split(b.pixel_values, per)
    mm = [dict(pixel_values=pvs[i], image_grid_thw=b.image_grid_thw[i:i + 1]) for i in range(3)]
    d = DataProto.from_dict(tensors=dict(input_ids=b.input_ids, attention_mask=b.attention_mask, response_mask=rm,
                                         position_ids=b.position_ids.transpose(0, 1).contiguous()),
                            non_tensors=dict(multi_modal_inputs=mm),
                            meta_info=dict(micro_batch_size=2, forward_args=dict(force_vit_image=True))).to("cuda")

    def forward_func_log_probs(data, logits):                      # the stock name -> fused LM head + log-prob kernel, no logits
        raise AssertionError("the fused path must not materialise logits")

    fused = st.forward_step(d, forward_func_log_probs)

    def custom(data, logits):                                      # any other forward_func gets real logits
        lp = st.op_compute_log_probs(logits, data.batch["input_ids"], data.batch["response_mask"])
        ent = st.op_compute_entropy(logits, data.batch["response_mask"])
        return lp, {"log_probs": lp, "entropy": ent}

    generic = st.forward_step(d, custom)
    for res in (fused, generic):
        assert res["log_probs"].shape == lp_ref.shape
        np.testing.assert_allclose(res["log_probs"].cpu().numpy(), lp_ref.numpy(), atol=6e-2)
        np.testing.assert_allclose(res["entropy"].cpu().numpy(), ent_ref.numpy(), atol=3e-2)
    np.testing.assert_allclose(fused["log_probs"].cpu().numpy(), generic["log_probs"].cpu().numpy(), atol=2e-2)
