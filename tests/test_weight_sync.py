"""CPU: train -> infer weight sync, receive side: the in-place kernel-layout update from HF-named tensors,
the bucket wire format of roll/utils/send_recv_utils.py, and a world_size-2 gloo broadcast through the strategy hooks
(worker_helper.py:64-116 behind vllm_strategy.py:258-271)."""
import os
import random
from types import SimpleNamespace

import pytest
import torch
import torch.multiprocessing as mp

from oracle import host_ref
from socioreasoner_b200.config import SocioReasonerConfig, TextConfig, VisionConfig
from socioreasoner_b200.roll_shim import B200InferStrategy
from socioreasoner_b200.roll_shim.send_recv import RecvBucketManager
from socioreasoner_b200.weights import PackedWeights, pack, random_state_dict, update_weights


def small_cfg(tied=True):
    # the fixed per-layer geometry (1280 / 2048) with 1 ViT block, 2 decoder layers, a small MLP and vocabulary
    return SocioReasonerConfig(text=TextConfig(num_hidden_layers=2, vocab_size=512, intermediate_size=352, tie_word_embeddings=tied),
                               vision=VisionConfig(depth=1, fullatt_block_indexes=(0,)))


def all_tensors(W: PackedWeights):
    out = {}
    for k, v in vars(W).items():
        if torch.is_tensor(v):
            out[k] = v
    for i, b in enumerate(W.vit_blocks):
        out.update({f"vit{i}.{k}": v for k, v in vars(b).items() if torch.is_tensor(v)})
    for i, l in enumerate(W.layers):
        out.update({f"dec{i}.{k}": v for k, v in vars(l).items() if torch.is_tensor(v)})
    return out


@pytest.mark.parametrize("tied", [True, False])
def test_update_weights_is_bit_exact_and_in_place(tied):
    cfg = small_cfg(tied)
    sd0 = random_state_dict(cfg, seed=1, device="cpu")
    sd1 = random_state_dict(cfg, seed=2, device="cpu")
    if not tied:
        sd0["lm_head.weight"] = torch.randn(512, 2048).bfloat16()
        sd1["lm_head.weight"] = torch.randn(512, 2048).bfloat16()
    W = pack(sd0, cfg, device="cpu")
    want = all_tensors(pack(sd1, cfg, device="cpu"))
    ptrs = {k: v.data_ptr() for k, v in all_tensors(W).items()}
    items = list(sd1.items())
    random.Random(0).shuffle(items)
    # half of the names in the transformers-5.x in-memory spelling
    items = [((("model.language_model." + k[len("model."):]) if (k.startswith("model.") and i % 2) else k), v) for i, (k, v) in enumerate(items)]
    done = update_weights(W, items, cfg)
    assert len(done) == len(items)
    got = all_tensors(W)
    assert got.keys() == want.keys()
    for k in want:
        assert got[k].data_ptr() == ptrs[k], k          # pointers stable: layer table / captured graphs stay valid
        assert torch.equal(got[k], want[k]), k
    assert (W.lm_head is W.embed) == tied
    with pytest.raises(KeyError):
        update_weights(W, [("model.layers.0.self_attn.rotary_emb.inv_freq", torch.zeros(4))], cfg)


def test_partial_update_touches_only_its_slice():
    cfg = small_cfg()
    sd0 = random_state_dict(cfg, seed=1, device="cpu")
    W = pack(sd0, cfg, device="cpu")
    before = {k: v.clone() for k, v in all_tensors(W).items()}
    new_k = torch.randn_like(sd0["model.layers.1.self_attn.k_proj.weight"].float()).bfloat16()
    update_weights(W, [("model.layers.1.self_attn.k_proj.weight", new_k)], cfg)
    after = all_tensors(W)
    changed = sorted(k for k in after if not torch.equal(after[k], before[k]))
    assert changed == ["dec1.qkv_w", "dec1.qkv_wp"]
    assert torch.equal(W.layers[1].qkv_w[2048:2304], new_k) and torch.equal(W.layers[1].qkv_w[:2048], before["dec1.qkv_w"][:2048])


@pytest.mark.parametrize("as_dict", [True, False])
def test_bucket_wire_format_roundtrip(as_dict):
    g = torch.Generator().manual_seed(0)
    named = [("a.weight", torch.randn(300, 70, generator=g).bfloat16()), ("b.bias", torch.randn(33, generator=g)),
             ("c.weight", torch.randn(1000, 129, generator=g).bfloat16()), ("d", torch.arange(7, dtype=torch.int64))]
    buckets = host_ref.send_buckets(named, bucket_size=50_000, as_dict=as_dict)   # c.weight spans 6 buckets
    assert len(buckets) > 5
    rx, got = RecvBucketManager(), {}
    for meta, buf in buckets:
        fin = rx.process_bucket(meta, buf)
        assert not (set(fin) & set(got))
        got.update(fin)
    rx.clear()
    assert list(got) == [n for n, _ in named]
    for n, t in named:
        assert got[n].dtype == t.dtype and got[n].shape == t.shape and torch.equal(got[n], t)


class RecordingModel:
    device = "cpu"

    def __init__(self):
        self.got = {}

    def load_weights(self, named):
        for n, t in named:
            assert n not in self.got
            self.got[n] = t.clone()


def _named(seed=3):
    g = torch.Generator().manual_seed(seed)
    return [("model.layers.0.mlp.down_proj.weight", torch.randn(256, 300, generator=g).bfloat16()),
            ("model.norm.weight", torch.randn(2048, generator=g)),
            ("visual.blocks.0.attn.qkv.weight", torch.randn(384, 128, generator=g).bfloat16())]


def _sync_worker(rank, port, q, done):
    from socioreasoner_b200.roll_shim import collective
    torch.manual_seed(0)
    named = _named()
    buckets = host_ref.send_buckets(named[:2], bucket_size=40_000)
    plan = {0: dict(group_name="model_update_test", master_addr="127.0.0.1", master_port=port, src_pp_rank=0,
                    tgt_devices=[dict(rank=5, device=dict(rank=0))])}
    if rank == 0:   # the train worker that holds the weights: rank 0 of the named group
        collective.init_collective_group(2, 0, backend="gloo", group_name="model_update_test", master_addr="127.0.0.1", master_port=port)
        collective.allreduce(torch.zeros(1), group_name="model_update_test")
        for meta, buf in buckets:
            collective.broadcast(buf, src_rank=0, group_name="model_update_test")
        collective.broadcast(named[2][1], src_rank=0, group_name="model_update_test")
        q.put(("sender", "ok"))
        done.wait(60)   # the store master leaves last: the receiver tears its side of the group down first
        collective.destroy_collective_group("model_update_test")
        return
    st = B200InferStrategy(SimpleNamespace(worker_config=None, rank=5))
    st.model = RecordingModel()
    st.setup_collective_group(plan, backend="gloo")
    assert st.model_update_comm_plan[0]["rank"] == 1 and st.model_update_comm_plan[0]["world_size"] == 2
    st.broadcast_bucket(7, {}, 10)            # unknown source pipeline rank: ignored (worker_helper.py:89-90)
    for meta, buf in buckets:
        st.broadcast_bucket(0, meta, buf.numel())
    st.broadcast_parameter(0, named[2][1].dtype, tuple(named[2][1].shape), named[2][0])
    st.recv_manager.clear()
    ok = all(torch.equal(st.model.got[n], t) for n, t in named) and list(st.model.got) == [n for n, _ in named]
    q.put(("receiver", "ok" if ok else "mismatch"))
    collective.destroy_collective_group("model_update_test")
    done.set()


def test_broadcast_bucket_and_parameter_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 150)
    done = ctx.Event()
    procs = [ctx.Process(target=_sync_worker, args=(r, port, q, done)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {"sender": "ok", "receiver": "ok"}


REF_PG_UTILS = "/root/reference/roll/utils/collective/pg_utils.py"


def _ref_pair_worker(rank, port, q, done):
    """rank 0 = the UNMODIFIED reference sender side (pg_utils.init_custom_process_group, a torch-only file loaded from the
    reference checkout), rank 1 = this repo's receiver strategy.  They must meet in the same named group."""
    import importlib.util
    import torch.distributed as dist
    named = _named(seed=9)
    plan = {0: dict(group_name="model_update_ref_pair", master_addr="127.0.0.1", master_port=port, src_pp_rank=0,
                    tgt_devices=[dict(rank=3, device=dict(rank=0))])}
    if rank == 0:
        spec = importlib.util.spec_from_file_location("ref_pg_utils", REF_PG_UTILS)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        # the reference picks the option keyword by a STRING compare (`str(torch.__version__) >= "2.6"`, pg_utils.py:59), which
        # misfires on torch 2.11 ("2.11" < "2.6"); it was written for torch 2.6 -- present that version string to it
        torch.__version__ = "2.6.0"
        pg = mod.init_custom_process_group(backend="gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=2, rank=0,
                                           group_name="model_update_ref_pair")      # collective.py:24-32 of the reference
        dist.all_reduce(torch.zeros(1), group=pg)                                  # megatron_strategy.py warm-up
        for _, t in named:
            dist.broadcast(t, src=0, group=pg)                                     # collective.broadcast of the reference
        q.put(("sender", "ok"))
        done.wait(60)
        return
    st = B200InferStrategy(SimpleNamespace(worker_config=None, rank=3))
    st.model = RecordingModel()
    st.setup_collective_group(plan, backend="gloo")
    for n, t in named:
        st.broadcast_parameter(0, t.dtype, tuple(t.shape), n)
    ok = all(torch.equal(st.model.got[n], t) for n, t in named)
    q.put(("receiver", "ok" if ok else "mismatch"))
    done.set()


@pytest.mark.skipif(not os.path.exists(REF_PG_UTILS), reason="reference checkout not present (GPU box)")
def test_group_rendezvous_with_reference_sender():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 150)
    done = ctx.Event()
    procs = [ctx.Process(target=_ref_pair_worker, args=(r, port, q, done)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {"sender": "ok", "receiver": "ok"}


def test_update_parameter_respects_ranks_in_worker():
    st = B200InferStrategy(SimpleNamespace(worker_config=None, rank=0))
    st.model = RecordingModel()
    w = torch.ones(3)
    st.update_parameter("model.norm.weight", w, ranks_in_worker=[1])       # not for this rank
    assert st.model.got == {}
    st.update_parameter("model.norm.weight", w, ranks_in_worker=[0])
    assert torch.equal(st.model.got["model.norm.weight"], w)
    meta, buf = host_ref.send_buckets([("x", torch.arange(10.0))], bucket_size=64)[0]
    st.update_parameter_in_bucket(meta, buf, ranks_in_worker=[0])
    assert torch.equal(st.model.got["x"], torch.arange(10.0))


def test_update_weights_accepts_every_key_of_an_hf_state_dict(oracle_tiny):
    """The names a real sender walks are the HF module's own state-dict keys (transformers 5.x spelling here: model.visual.*,
    model.language_model.*, tied lm_head): every key must be consumed, and the result must equal a fresh pack."""
    from oracle import hf_oracle as O
    cfg = SocioReasonerConfig.from_hf(oracle_tiny.config)
    sd_a = {k: v.detach().clone() for k, v in oracle_tiny.state_dict().items()}
    other = O.build_model("tiny", seed=43)
    sd_b = {k: v.detach().clone() for k, v in other.state_dict().items()}
    assert sd_a.keys() == sd_b.keys() and any(k.startswith("model.visual.") for k in sd_a)
    W = pack(sd_a, cfg, device="cpu")
    done = update_weights(W, list(sd_b.items()), cfg)
    assert sorted(done) == sorted(sd_b)
    want, got = all_tensors(pack(sd_b, cfg, device="cpu")), all_tensors(W)
    for k in want:
        assert torch.equal(got[k], want[k]), k
