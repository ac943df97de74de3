"""GPU (-m gpu, needs >= 2 GPUs; skipped otherwise): the train -> infer weight broadcast over NCCL through the named group the
strategies build (roll/third_party/vllm/worker_helper.py:64-116 behind vllm_strategy.py:258-271): rank 0 = the sender (the train
worker of the reference), rank 1 = ``B200InferStrategy`` receiving whole parameters and int8 buckets into a real model's
kernel-layout weights, bit-exact; plus the time of a 1 GiB bucketed broadcast against the NVLink rate."""
import os
import time
from types import SimpleNamespace

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, port, q):
    from oracle import host_ref
    from socioreasoner_b200.config import SocioReasonerConfig, TextConfig, VisionConfig
    from socioreasoner_b200.roll_shim import B200InferStrategy, collective
    from socioreasoner_b200.weights import pack, random_state_dict

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    cfg = SocioReasonerConfig(text=TextConfig(num_hidden_layers=2, vocab_size=512, intermediate_size=352),
                              vision=VisionConfig(depth=1, fullatt_block_indexes=(0,)))
    sd_new = random_state_dict(cfg, seed=2, device="cpu")
    names = list(sd_new)
    big = torch.empty(1 << 28, dtype=torch.int8, device=dev)       # 256 MiB bucket, broadcast 4 times
    plan = {0: dict(group_name="model_update_nccl", master_addr="127.0.0.1", master_port=port, src_pp_rank=0,
                    tgt_devices=[dict(rank=9, device=dict(rank=0))])}
    if rank == 0:
        collective.init_collective_group(2, 0, backend="nccl", group_name="model_update_nccl", master_addr="127.0.0.1", master_port=port)
        collective.allreduce(torch.zeros(1, device=dev), group_name="model_update_nccl")
        half = len(names) // 2
        for n in names[:half]:                                       # whole-parameter broadcasts
            collective.broadcast(sd_new[n].to(dev), src_rank=0, group_name="model_update_nccl")
        for meta, buf in host_ref.send_buckets([(n, sd_new[n]) for n in names[half:]], bucket_size=1 << 20):
            collective.broadcast(buf.to(dev), src_rank=0, group_name="model_update_nccl")
        torch.cuda.synchronize()
        for _ in range(5):
            collective.broadcast(big, src_rank=0, group_name="model_update_nccl")
        torch.cuda.synchronize()
        q.put(("sender", "ok"))
        return
    from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration as Model
    from socioreasoner_b200.weights import PackedWeights
    model = Model(cfg, pack(random_state_dict(cfg, seed=1, device="cpu"), cfg, dev), dev)
    st = B200InferStrategy(SimpleNamespace(worker_config=None, rank=9))
    st.model = model
    st.setup_collective_group(plan, backend="nccl")
    half = len(names) // 2
    for n in names[:half]:
        st.broadcast_parameter(0, sd_new[n].dtype, tuple(sd_new[n].shape), n)
    for meta, buf in host_ref.send_buckets([(n, sd_new[n]) for n in names[half:]], bucket_size=1 << 20):
        st.broadcast_bucket(0, meta, buf.numel())
    st.recv_manager.clear()
    want = pack(sd_new, cfg, dev)
    ok = all(torch.equal(getattr(model.engine.w, k), getattr(want, k)) for k in ("embed", "final_norm", "patch_w", "merger_w0")) and \
        all(torch.equal(a.qkv_wp, b.qkv_wp) and torch.equal(a.down_w, b.down_w) for a, b in zip(model.engine.w.layers, want.layers))
    torch.cuda.synchronize()
    collective.broadcast(big, src_rank=0, group_name="model_update_nccl")   # warm-up
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(4):
        collective.broadcast(big, src_rank=0, group_name="model_update_nccl")
    torch.cuda.synchronize()
    gbs = 4 * big.numel() / (time.time() - t0) / 1e9
    q.put(("receiver", "ok" if ok else "mismatch", round(gbs, 1)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_weight_sync_over_nccl_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = {r[0]: r[1:] for r in res}
    assert d["sender"][0] == "ok" and d["receiver"][0] == "ok", d
    print(f"\n1 GiB bucketed broadcast over the named NCCL group: {d['receiver'][1]} GB/s (NVLink 5: 900 GB/s per direction)")
    assert d["receiver"][1] > 50
