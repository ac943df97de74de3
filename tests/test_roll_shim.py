"""CPU: the mirrored reference surface (DataProto subset, strategy micro-batch loop, factory) with a fake model, and the
world_size-2 (gloo) data-parallel shard + gather logic."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import host_ref
from socioreasoner_b200.roll_shim import B200InferStrategy, B200SegStrategy, DataProto, collate_fn_to_dict_list, create_strategy
from socioreasoner_b200.rollout import DataParallelRollout, shard_indices


def make_proto(n=7, S=5):
    ids = torch.arange(n * S).view(n, S)
    mm = [dict(pixel_values=torch.full((4 * (i + 1), 1176), float(i)), image_grid_thw=torch.tensor([[1, 2, 2 * (i + 1)]])) for i in range(n)]
    return DataProto.from_dict(
        tensors=dict(input_ids=ids, attention_mask=torch.ones_like(ids), position_ids=torch.arange(n * 3 * S).view(n, 3, S),
                     response_mask=torch.ones_like(ids)),
        non_tensors=dict(multi_modal_inputs=mm), meta_info=dict(micro_batch_size=2))


def test_dataproto_chunk_concat_follow_array_split():
    d = make_proto(7)
    for chunks in (1, 2, 3, 4, 7):
        parts = d.chunk(chunks)
        assert [len(p) for p in parts] == [len(c) for c in np.array_split(np.arange(7), chunks)]
        back = DataProto.concat(parts)
        assert torch.equal(back.batch["input_ids"], d.batch["input_ids"])
        assert [m["image_grid_thw"].tolist() for m in back.non_tensor_batch["multi_modal_inputs"]] == \
               [m["image_grid_thw"].tolist() for m in d.non_tensor_batch["multi_modal_inputs"]]
        assert all(p.meta_info is d.meta_info for p in parts)
    with pytest.raises(AssertionError):
        d.chunk(8)


def test_collate_fn_to_dict_list():
    out = collate_fn_to_dict_list([{"a": torch.ones(2, 3), "m": np.zeros(2)}, {"a": torch.zeros(1, 3), "m": np.ones(2)}])
    assert out["a"].shape == (3, 3) and out["m"].dtype == object and len(out["m"]) == 2


class FakeModel:
    """Records what the strategy passes (the contract of hf_strategy.py:84-90) and returns logits that encode the ids."""
    def __init__(self):
        self.calls = []

    def __call__(self, input_ids, attention_mask, position_ids, use_cache, **kw):
        self.calls.append(dict(ids=input_ids.clone(), pos_shape=tuple(position_ids.shape), use_cache=use_cache,
                               pv_rows=kw["pixel_values"].shape[0], grid=kw["image_grid_thw"].clone(), extra=sorted(kw)))
        V = 64
        logits = torch.zeros(*input_ids.shape, V)
        logits.scatter_(-1, (input_ids % V).unsqueeze(-1), 5.0)
        return SimpleNamespace(logits=logits)

    def generate(self, input_ids, attention_mask, use_cache, **kw):
        n = kw["max_new_tokens"] - (1 if int(input_ids[0, 0]) % 2 else 0)  # ragged outputs across micro-batches
        return torch.cat([input_ids, torch.full((input_ids.shape[0], n), 9)], 1)


def worker(name):
    return SimpleNamespace(worker_config=SimpleNamespace(strategy_args=SimpleNamespace(strategy_name=name), model_args=None),
                           rank_info=SimpleNamespace(dp_rank=None, dp_size=None))


def test_forward_step_micro_batching_and_multimodal_concat():
    st = create_strategy(worker("hf_infer"))
    assert isinstance(st, B200InferStrategy)
    st.model = FakeModel()
    d = make_proto(7)

    def forward_func(data, logits):
        lp = host_ref.op_compute_log_probs(logits, data.batch["input_ids"], data.batch["response_mask"])
        return lp, {"log_probs": lp.clone(), "tag": "x"}

    res = st.forward_step(d, forward_func)
    # 7 // 2 = 3 micro-batches with array_split sizes 3,2,2; results concatenated in order
    assert [c["ids"].shape[0] for c in st.model.calls] == [3, 2, 2]
    assert res["log_probs"].shape == (7, 4) and len(res["tag"]) == 3
    ref = host_ref.op_compute_log_probs(FakeModel()(d.batch["input_ids"], None, d.batch["position_ids"], False,
                                                    pixel_values=torch.zeros(1, 1), image_grid_thw=torch.zeros(1, 3)).logits,
                                        d.batch["input_ids"], d.batch["response_mask"])
    assert torch.allclose(res["log_probs"], ref)
    c0 = st.model.calls[0]
    assert c0["pos_shape"] == (3, 3, 5) and c0["use_cache"] is False          # (bsz,3,S) -> (3,bsz,S)
    assert c0["pv_rows"] == 4 + 8 + 12 and c0["grid"].shape == (3, 3)          # per-sample mm inputs concatenated
    assert c0["extra"] == ["image_grid_thw", "pixel_values"]


def test_forward_step_takes_the_fused_log_prob_path_for_the_stock_forward_func():
    """forward_step(forward_func_log_probs) (base_worker.py:226-246) -> model.compute_log_probs (no logits); any other
    forward_func keeps the generic logits path.  Same keys, shapes and micro-batch collation either way."""
    class FusedModel(FakeModel):
        def __init__(self):
            super().__init__()
            self.fused_calls = []

        def compute_log_probs(self, input_ids, attention_mask, position_ids, response_mask, pixel_values=None,
                              image_grid_thw=None, **kw):
            assert "force_vit_image" not in kw
            self.fused_calls.append(input_ids.shape[0])
            logits = FakeModel()(input_ids, None, position_ids, False, pixel_values=pixel_values, image_grid_thw=image_grid_thw).logits
            return (host_ref.op_compute_log_probs(logits, input_ids, response_mask),
                    host_ref.op_compute_entropy(logits, response_mask))

    st = create_strategy(worker("hf_infer"))
    st.model = FusedModel()
    d = make_proto(7)
    d.batch["response_mask"][:, :2] = 0
    d.meta_info["forward_args"] = {"force_vit_image": True}

    def forward_func_log_probs(data, logits):  # the stock reduction, by name
        lp = host_ref.op_compute_log_probs(logits, data.batch["input_ids"], data.batch["response_mask"])
        ent = host_ref.op_compute_entropy(logits, data.batch["response_mask"])
        return lp, {"log_probs": lp.clone(), "entropy": ent.clone()}

    fused = st.forward_step(d, forward_func_log_probs)
    assert st.model.fused_calls == [3, 2, 2] and st.model.calls == []
    st.fuse_log_probs = False
    plain = st.forward_step(d, forward_func_log_probs)
    assert [c["ids"].shape[0] for c in st.model.calls] == [3, 2, 2]
    for k in ("log_probs", "entropy"):
        assert fused[k].shape == plain[k].shape == (7, 4)
        assert torch.allclose(fused[k], plain[k])
    st.fuse_log_probs = True
    st.model.calls.clear()
    st.forward_step(d, lambda data, logits: (logits.sum(), {"s": logits.sum(-1)}))  # custom forward_func: generic path
    assert len(st.model.calls) == 3


def test_postprocessed_generate_layout_feeds_the_fused_log_prob_path():
    """generate -> postprocess_generate (functionals.py:768-872: right padding, response_mask, continued M-RoPE ids) ->
    compute_log_probs: the rows the fused path scores are exactly the response tokens of every sequence, labelled with them."""
    from socioreasoner_b200 import plan
    pad, eos, S_p, T, seq_len, n = 0, 9, 7, 5, 14, 2
    ids = torch.tensor([[0, 0, 0, 3, 4, 5, 6], [0, 1, 2, 3, 4, 5, 6], [7, 1, 2, 3, 4, 5, 6]])
    am = (ids != 0).long()
    pos = torch.cumsum(am, -1).sub(1).clamp(min=0)[:, None, :].repeat(1, 3, 1)               # text-only M-RoPE ids
    gen = torch.tensor([[11, 12, eos, pad, pad], [13, eos, pad, pad, pad], [14, 15, 16, 17, 18], [19, 20, 21, eos, pad],
                        [22, 23, 24, 25, eos], [26, eos, pad, pad, pad]])
    output = torch.cat([ids.repeat_interleave(n, 0), gen], 1)                                  # what strategy.generate returns
    d = host_ref.postprocess_generate_ref(ids, am, pos, output, n, seq_len, pad)
    assert d["input_ids"].shape == (6, seq_len) and d["position_ids"].shape == (6, 3, seq_len)
    for i in range(6):
        p_tok = ids[i // n][am[i // n].bool()].tolist()
        r_tok = [t for t in gen[i].tolist() if t != pad]
        valid = len(p_tok) + len(r_tok)
        assert d["input_ids"][i, :valid].tolist() == p_tok + r_tok                              # right padded, nothing lost
        assert d["attention_mask"][i].tolist() == [1] * valid + [0] * (seq_len - valid)
        assert d["response_mask"][i].nonzero().flatten().tolist() == list(range(len(p_tok), valid))
        assert d["position_ids"][i, 0, :valid].tolist() == list(range(valid))                   # positions continue after the prompt
    flat = np.nonzero(d["attention_mask"].numpy().reshape(-1))[0]
    rows, labels, out = plan.logprob_rows(d["input_ids"].numpy(), d["response_mask"].numpy(), flat)
    want_labels = [t for i in range(6) for t in gen[i].tolist() if t != pad]
    assert labels.tolist() == want_labels                                                       # every response token, once, in order
    packed_ids = d["input_ids"].numpy().reshape(-1)[flat]
    assert (packed_ids[rows + 1] == labels).all()                                               # row s predicts token s + 1
    b, s_ = out // (seq_len - 1), out % (seq_len - 1)
    assert (d["response_mask"].numpy()[b, s_ + 1] == 1).all()


def test_generate_pads_ragged_micro_batches():
    st = create_strategy(worker("vllm"))
    st.model = FakeModel()
    d = make_proto(4)
    d.meta_info["micro_batch_size"] = 2
    out = st.generate(d, dict(max_new_tokens=3, pad_token_id=0))
    assert out.shape == (4, 5 + 3)
    assert (out[:, :5] == d.batch["input_ids"]).all()


def test_factory_names_and_seg_empty_prompt():
    assert isinstance(create_strategy(worker("seg_infer")), B200SegStrategy)
    with pytest.raises(ValueError, match="Unknown strategy name"):
        create_strategy(worker("deepspeed_train"))
    seg = create_strategy(worker("seg_infer"))
    seg.predictor = lambda image, vp: (_ for _ in ()).throw(RuntimeError("must not be called"))
    d = DataProto.from_dict(tensors=dict(x=torch.zeros(2, 1)), non_tensors=dict(seg_image=[None, None], visual_prompt=[[], []]))
    out = seg.segment(d)
    assert out["mask"].shape == (2,) and out["mask"][0].shape == (768, 768) and out["mask"][0].sum() == 0


def test_offload_hooks_are_noops():
    st = B200InferStrategy(worker("b200_infer"))
    assert st.load_states(include=None, non_blocking=True) is None and st.offload_states() is None


# ---------------------------------------------------------------------------------------------- world_size 2, gloo
def _dp_worker(rank, world, port, total, T, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = shard_indices(total, world)
        mine = shards[rank]
        b_max = max(len(s) for s in shards)
        ro = DataParallelRollout(None, max_new_tokens=T, out_size=8)
        masks = ro.slot_masks(b_max, device="cpu")
        for j, i in enumerate(mine):
            masks[j].fill_(int(i) + 1)                              # "epilogue" writes straight into the gather slot
        toks = torch.stack([torch.arange(T) + 100 * int(i) for i in mine]) if len(mine) else torch.zeros(0, T, dtype=torch.int64)
        ro.pack_slot(b_max, toks, torch.tensor([float(i) / 10 for i in mine]))
        res = ro.gather(b_max)
        # driver-side collect: destination rank gets per-rank host views (no concatenation), the others None
        pend = ro.gather_to_host_async(b_max, dst_rank=1)
        assert (pend is None) == (rank != 1)
        if pend is not None:
            ht, hm, hi = pend.wait()
            assert [len(t) for t in ht] == [len(s) for s in shards]
            assert torch.equal(torch.cat(ht), res.tokens) and torch.equal(torch.cat(hm), res.masks) and torch.equal(torch.cat(hi), res.iou)
        # plain numpy through the queue: torch tensors travel as shared-memory handles that die with the producer process
        out_q.put((rank, res.tokens.numpy().copy(), res.masks.numpy().copy(), res.iou.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 4])
def test_dp_shard_and_gather_world2_gloo(total):
    world, T = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + total + (os.getpid() % 200) * 10
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, total, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, toks, masks, iou in results:
        toks, masks, iou = torch.from_numpy(toks), torch.from_numpy(masks), torch.from_numpy(iou)
        # every rank sees the whole batch in rank order == np.array_split order == the reference's DataProto.concat order
        assert toks.shape == (total, T) and masks.shape == (total, 8, 8)
        assert toks[:, 0].tolist() == [100 * i for i in range(total)]
        assert masks[:, 0, 0].tolist() == [i + 1 for i in range(total)]
        assert torch.allclose(iou, torch.arange(total).float() / 10)


# ---- initialize + the vLLM input form (what the unchanged entry script sends) ----------------------------------------------
def test_initialize_sets_tokenizer_processor_model_and_rank_info(tmp_path, oracle_tiny):
    """hf_strategy.py:34-47 / base_worker.py:46-49,129-131: after initialize the worker reads strategy.tokenizer (eos / pad ids)."""
    import ckpt_util
    path = ckpt_util.write_checkpoint(str(tmp_path / "ckpt"), oracle_tiny)
    w = worker("vllm")
    w.worker_config.model_args = SimpleNamespace(model_name_or_path=path)
    st = create_strategy(w)
    seen = {}

    def provider(tokenizer, model_args, is_trainable=False):
        seen.update(tokenizer=tokenizer, path=model_args.model_name_or_path, trainable=is_trainable)
        return FakeModel()

    st.initialize(provider)
    assert st.tokenizer is seen["tokenizer"] and st.tokenizer is not None and seen["path"] == path and seen["trainable"] is False
    assert st.tokenizer.eos_token_id == 7002 and st.tokenizer.pad_token_id == 7000 and st.tokenizer.padding_side == "left"
    assert st.processor is not None and isinstance(st.model, FakeModel)
    assert w.rank_info.dp_rank == 0 and w.rank_info.dp_size == 1


def test_generate_accepts_the_vllm_input_form(monkeypatch):
    """vllm_strategy.py:114-141: prompts = {prompt_token_ids, multi_modal_data: {image: [PIL, ...]}}; output = prompt + responses
    right-padded to the longest response of the WHOLE batch (gather_outputs_to_pad_tensor + concatenate_input_and_output)."""
    from PIL import Image
    st = create_strategy(worker("vllm"))
    calls = []

    class GenModel:
        device = "cpu"

        def generate(self, input_ids, attention_mask, position_ids, pixel_values, image_grid_thw, max_new_tokens, eos_token_id,
                     pad_token_id, do_sample, temperature, top_p, top_k, num_return_sequences, **kw):
            calls.append(dict(b=input_ids.shape[0], rows=pixel_values.shape[0], grid=image_grid_thw.tolist(), pos=tuple(position_ids.shape),
                              pos_t=position_ids.clone(),
                              sample=(do_sample, temperature, top_p, top_k, num_return_sequences)))
            new = torch.full((input_ids.shape[0], max_new_tokens), pad_token_id)
            for i in range(input_ids.shape[0]):
                k = 2 + int(input_ids[i, -1]) % 3                      # ragged response lengths, stop token included
                new[i, :k] = 11
                new[i, k - 1] = eos_token_id[0]
            return torch.cat([input_ids, new], 1)

    st.model = GenModel()
    monkeypatch.setattr(type(st), "_image_preprocessor", lambda self: (lambda images, out_dtype=None: (
        torch.zeros(sum((im.shape[0] // 14) * (im.shape[1] // 14) for im in images), 1176),
        torch.tensor([[1, im.shape[0] // 14, im.shape[1] // 14] for im in images]))))
    n = 5
    ids = torch.arange(n * 6).view(n, 6)
    mm = [{"prompt_token_ids": [1, 2, 3], "multi_modal_data": {"image": [Image.new("RGB", (56, 28 * (1 + i % 2))), Image.new("RGB", (28, 28))]}}
          for i in range(n)]
    pos_b3s = torch.arange(n * 3 * 6).view(n, 3, 6)
    d = DataProto.from_dict(tensors=dict(input_ids=ids, attention_mask=torch.ones_like(ids), position_ids=pos_b3s),
                            non_tensors=dict(multi_modal_data=mm), meta_info=dict(micro_batch_size=2))
    gc = dict(max_new_tokens=8, temperature=1.0, top_p=0.8, top_k=100, num_beams=1, repetition_penalty=1.0, num_return_sequences=1,
              eos_token_id=[99, 98], pad_token_id=0)
    out = st.generate(d, gc)
    assert [c["b"] for c in calls] == [3, 2] and calls[0]["pos"] == (3, 3, 6)               # micro-batches; (b,3,S) -> (3,b,S)
    assert torch.equal(calls[0]["pos_t"], pos_b3s[:3].transpose(0, 1)) and torch.equal(calls[1]["pos_t"], pos_b3s[3:].transpose(0, 1))
    assert calls[0]["grid"] == [[1, 2, 4], [1, 2, 2], [1, 4, 4], [1, 2, 2], [1, 2, 4], [1, 2, 2]]   # two images per prompt, in order
    assert calls[0]["sample"] == (True, 1.0, 0.8, 100, 1)
    assert out.shape == (n, 6 + 4) and torch.equal(out[:, :6], ids)                        # longest response: 2 + 2 tokens
    lens = [2 + int(ids[i, -1]) % 3 for i in range(n)]
    for i, k in enumerate(lens):
        assert out[i, 6 + k - 1] == 99 and (out[i, 6 + k:] == 0).all() and (out[i, 6:6 + k - 1] == 11).all()
