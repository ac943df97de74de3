"""Pin the oracle's restatements of the reference's IN-TREE algorithms on outputs of the reference itself.

`import roll` / `import mcore_adapter` fail in this image (ray, tensordict, megatron, sam2 ... are absent -- SURVEY.md 8c), but
the algorithms on the hot path are plain functions / methods.  This script reads their SOURCE from /root/reference with `ast`,
executes exactly those definitions (unmodified text) in a namespace that holds only torch / numpy / cv2 and small stand-ins for
the container types they construct (TensorDict, DataProto), runs them on seeded inputs and stores inputs + outputs under
tests/golden/reference_intree.pt.  tests/test_reference_fixtures.py then checks oracle/host_ref.py and the product's host
logic against that file -- on any machine, without /root/reference.

Functions executed (file:line in the reference):
  mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py  Qwen2_5_VLModel.get_rope_index      :319-441
  (same file)  rotate_half :14, apply_multimodal_rotary_pos_emb :21-41, Qwen2_5_VLRotaryEmbedding.forward :62-75
  roll/utils/functionals.py   log_probs_from_logits :193, entropy_from_logits :200, get_pad_mask :301, pad_to_length :351,
                              postprocess_generate :768-872, get_dist_info_from_comm_plan :875
  roll/distributed/strategy/strategy.py      InferenceStrategy.op_compute_log_probs / op_compute_entropy   :123-138
  roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py   compute_giou                                    :45-58
  roll/distributed/strategy/seg_strategy.py  SegInferStrategy.segment (pick / OR / resize around a stub predictor) :26-72
  roll/datasets/collator.py                  collate_fn_to_dict_list                                       :13-40
  roll/utils/send_recv_utils.py              TensorBucket, SendBucketManager, RecvBucketManager            :64-163

Usage (only where /root/reference exists): python tools/make_reference_fixtures.py
"""
import ast
import os
import sys
import textwrap
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("SR_REFERENCE", "/root/reference")


def extract(path, names):
    """Source text of top-level functions / classes, or `Class.method` (dedented), exactly as written in the reference."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
        if isinstance(node, ast.ClassDef):
            for sub in node.body:
                key = f"{node.name}.{sub.name}" if isinstance(sub, ast.FunctionDef) else None
                if key in names:
                    out[key] = textwrap.dedent(ast.get_source_segment(src, sub, padded=True))
    missing = set(names) - set(out)
    assert not missing, (path, missing)
    return out


def run_defs(defs, ns):
    for text in defs.values():
        exec(compile(text, "<reference>", "exec"), ns)
    return ns


def base_ns():
    import torch.nn.functional as F
    from typing import Dict, List, Optional, Tuple, Union
    return dict(torch=torch, np=np, F=F, Optional=Optional, Tuple=Tuple, List=List, Dict=Dict, Union=Union)


# ---------------------------------------------------------------------------------------------------------------------
class TensorDict(dict):  # stand-in for tensordict.TensorDict as postprocess_generate constructs it
    def __init__(self, source, batch_size=None):
        super().__init__(source)
        self.batch_size = batch_size


class DataProtoStub:  # stand-in for roll.distributed.scheduler.protocol.DataProto (fields the functions touch)
    def __init__(self, batch=None, non_tensor_batch=None, meta_info=None):
        self.batch, self.non_tensor_batch, self.meta_info = batch, non_tensor_batch or {}, meta_info or {}

    def chunk(self, chunks):
        n = self.batch.batch_size[0]
        assert chunks == n
        return [DataProtoStub(batch=SimpleNamespace(batch_size=[1]), non_tensor_batch={k: v[i:i + 1] for k, v in self.non_tensor_batch.items()})
                for i in range(n)]


def main():
    from oracle import hf_oracle as O
    from oracle import host_ref
    fx = {}

    # ---- get_rope_index (in-tree copy = authoritative) ----
    d = extract("mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py", ["Qwen2_5_VLModel.get_rope_index"])
    ns = run_defs(d, base_ns())
    cfg = SimpleNamespace(merge_size=2, image_token_id=151655, video_token_id=151656, vision_start_token_id=151652, tokens_per_second=2)
    self_stub = SimpleNamespace(config=cfg)
    g = torch.Generator().manual_seed(0)

    def prompt(grids, n_text, tail=3):
        ids = [151644, 872, 198]
        for (t, h, w) in grids:
            ids += [151652] + [151655] * (t * h * w // 4) + [151653]
            ids += torch.randint(0, 151643, (n_text,), generator=g).tolist()
        return ids + [151645, 198][:tail]

    cases = []
    for grids_per_sample, n_text in ([[[(1, 4, 6)]], 5], [[[(1, 4, 4), (1, 6, 4)], [(1, 8, 8)]], 7], [[[], [(1, 2, 2)]], 4], [[[], []], 6]):
        rows = [prompt(gs, n_text) for gs in grids_per_sample]
        S = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), S, dtype=torch.long)
        am = torch.zeros_like(ids)
        for i, r in enumerate(rows):           # left padding, as the reference collator produces
            ids[i, S - len(r):] = torch.tensor(r)
            am[i, S - len(r):] = 1
        flat = [gg for gs in grids_per_sample for gg in gs]
        grid = torch.tensor(flat, dtype=torch.long) if flat else None
        pos, deltas = ns["get_rope_index"](self_stub, ids, image_grid_thw=grid, attention_mask=am)
        cases.append(dict(input_ids=ids, attention_mask=am, image_grid_thw=grid, position_ids=pos, rope_deltas=deltas))
    fx["rope_index"] = cases

    # ---- M-RoPE table + apply (in-tree copies, mcore ...:14-75) ----
    d = extract("mcore_adapter/src/mcore_adapter/models/qwen2_5_vl/modeling_qwen2_5_vl.py",
                ["rotate_half", "apply_multimodal_rotary_pos_emb", "Qwen2_5_VLRotaryEmbedding.forward"])
    rns = run_defs(d, base_ns())
    inv_freq = 1.0 / (1000000.0 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    Bm, Sm = 2, 11
    pos3 = torch.randint(0, 4000, (3, Bm, Sm), generator=g)
    emb = rns["forward"](SimpleNamespace(inv_freq=inv_freq), torch.zeros(1), pos3)               # [3, B, S, 128]
    q = torch.randn(Sm, Bm, 4, 128, generator=g)
    k = torch.randn(Sm, Bm, 2, 128, generator=g)
    qe, ke = rns["apply_multimodal_rotary_pos_emb"](q, k, emb.cos(), emb.sin(), mrope_section=[16, 24, 24])
    fx["mrope"] = dict(position_ids=pos3, inv_freq=inv_freq, mrope_section=[16, 24, 24], emb=emb, q=q, k=k, q_embed=qe, k_embed=ke)

    # ---- functionals + strategy ops ----
    d = extract("roll/utils/functionals.py", ["log_probs_from_logits", "entropy_from_logits", "get_pad_mask", "pad_to_length",
                                              "postprocess_generate", "get_dist_info_from_comm_plan"])
    fns = run_defs(d, base_ns())
    d2 = extract("roll/distributed/strategy/strategy.py", ["InferenceStrategy.op_compute_log_probs", "InferenceStrategy.op_compute_entropy"])
    fns2 = run_defs(d2, dict(base_ns(), log_probs_from_logits=fns["log_probs_from_logits"], entropy_from_logits=fns["entropy_from_logits"]))
    logits = torch.randn(3, 9, 37, generator=g) * 3
    ids = torch.randint(0, 37, (3, 9), generator=g)
    rm = torch.tensor([[0, 0, 0, 1, 1, 1, 1, 0, 0], [0, 1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 0, 0, 0, 0]])
    fx["logprob"] = dict(logits=logits, input_ids=ids, response_mask=rm,
                         log_probs_from_logits=fns["log_probs_from_logits"](logits[:, :-1], ids[:, 1:]),
                         entropy_from_logits=fns["entropy_from_logits"](logits),
                         op_compute_log_probs=fns2["op_compute_log_probs"](None, logits, ids, rm),
                         op_compute_entropy=fns2["op_compute_entropy"](None, logits, rm))

    # ---- postprocess_generate (M-RoPE branch), with stand-in containers ----
    import types
    fake_protocol = types.ModuleType("roll.distributed.scheduler.protocol")
    fake_protocol.DataProto = DataProtoStub
    for name in ("roll", "roll.distributed", "roll.distributed.scheduler"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["roll.distributed.scheduler.protocol"] = fake_protocol
    fns["TensorDict"] = TensorDict
    pad, eos, n, seq_len = 0, 9, 2, 14
    p_ids = torch.tensor([[0, 0, 0, 3, 4, 5, 6], [0, 1, 2, 3, 4, 5, 6], [7, 1, 2, 3, 4, 5, 6]])
    p_am = (p_ids != 0).long()
    p_pos = torch.cumsum(p_am, -1).sub(1).clamp(min=0)[:, None, :].repeat(1, 3, 1)
    p_pos[0, 1] += 2; p_pos[0, 2] += 5              # distinct axes, as after an image span
    gen = torch.tensor([[11, 12, eos, pad, pad], [13, eos, pad, pad, pad], [14, 15, 16, 17, 18], [19, 20, 21, eos, pad],
                        [22, 23, 24, 25, eos], [26, eos, pad, pad, pad]])
    output = torch.cat([p_ids.repeat_interleave(n, 0), gen], 1)
    prompts = DataProtoStub(batch={"input_ids": p_ids.clone(), "attention_mask": p_am.clone(), "position_ids": p_pos.clone()})
    res = fns["postprocess_generate"](prompts, output.clone(), n, seq_len, eos, pad)
    fx["postprocess_generate"] = dict(input_ids=p_ids, attention_mask=p_am, position_ids=p_pos, output=output, num_return_sequences=n,
                                      sequence_length=seq_len, pad_token_id=pad, result={k: v.clone() for k, v in res.batch.items()})

    # ---- comm plan lookup ----
    plan = {0: dict(group_name="g0", master_addr="127.0.0.1", master_port=1, src_pp_rank=0,
                    tgt_devices=[dict(rank=4, device=dict(rank=0)), dict(rank=5, device=dict(rank=0)), dict(rank=5, device=dict(rank=1))])}
    fx["comm_plan"] = dict(plan=plan, lookups=[((r, w), fns["get_dist_info_from_comm_plan"](plan, r, w)[0]) for r, w in
                                               [(4, 0), (5, 0), (5, 1), (6, 0)]])

    # ---- compute_giou ----
    d = extract("roll/pipeline/rlvr/rlvr_socioseg_vlm_pipeline_infer.py", ["compute_giou"])
    giou = run_defs(d, base_ns())["compute_giou"]
    rs = np.random.RandomState(0)
    masks = [(rs.rand(64, 64) < p).astype(np.uint8) * v for p, v in ((0.3, 1), (0.5, 255), (0.0, 1), (1.0, 1))]
    pairs = [(0, 1), (1, 1), (2, 2), (0, 2), (3, 0)]
    fx["giou"] = dict(masks=masks, pairs=pairs, values=[float(giou(masks[a], masks[b])) for a, b in pairs])

    # ---- seg epilogue: the real segment() around a stub predictor ----
    import cv2
    from PIL import Image
    d = extract("roll/datasets/collator.py", ["collate_fn_to_dict_list"])
    coll_ns = run_defs(d, dict(base_ns(), defaultdict=__import__("collections").defaultdict))
    d = extract("roll/distributed/strategy/seg_strategy.py", ["SegInferStrategy.segment"])
    seg = run_defs(d, dict(base_ns(), cv2=cv2, DataProto=DataProtoStub, collate_fn_to_dict_list=coll_ns["collate_fn_to_dict_list"]))["segment"]

    class StubPredictor:  # SAM2ImagePredictor surface used by segment(): set_image, predict(**prompt) -> (masks [3,756,756], scores [3], _)
        def __init__(self, per_object):
            self.per_object, self.i = per_object, 0

        def set_image(self, image):
            assert image.size == (756, 756)
            self.i = 0

        def predict(self, **prompt):
            item = self.per_object[self.i]
            self.i += 1
            if item is None:
                raise RuntimeError("predictor failure")   # the reference skips the object
            return item[0], item[1], None

    def obj_array(x):      # a length-1 object array holding x itself (np.array would unpack PIL images / lists)
        a = np.empty(1, dtype=object)
        a[0] = x
        return a

    seg_cases = []
    for tile, n_obj, fail in ((0, 3, None), (1, 1, None), (2, 4, 1)):
        ml, ms, _gt = O.synth_mask_inputs(tile, n_obj)
        up = torch.nn.functional.interpolate(ml, size=(756, 756), mode="bilinear", align_corners=False) > 0   # SAM2's tail (restated)
        per_obj = [None if j == fail else (up[j].numpy(), ms[j].numpy()) for j in range(n_obj)]
        strat = SimpleNamespace(model=StubPredictor(per_obj))
        batch = DataProtoStub(batch=SimpleNamespace(batch_size=[1]),
                              non_tensor_batch={"seg_image": obj_array(Image.new("RGB", (1024, 1024))),
                                                "visual_prompt": obj_array([{"box": [0, 0, 1, 1]}] * n_obj)})
        out = seg(strat, batch)
        keep = [j for j in range(n_obj) if j != fail]
        # inputs are re-derivable from the seeds (O.synth_mask_inputs(tile, n_obj), minus the failing object); masks bit-packed
        seg_cases.append(dict(tile=tile, n_obj=n_obj, fail=fail, keep=keep, mask_bits=np.packbits(np.asarray(out["mask"][0]).astype(bool)),
                              mask_max=int(np.asarray(out["mask"][0]).max())))
    empty = DataProtoStub(batch=SimpleNamespace(batch_size=[1]),
                          non_tensor_batch={"seg_image": obj_array(Image.new("RGB", (64, 64))), "visual_prompt": obj_array([])})
    m_empty = np.asarray(seg(SimpleNamespace(model=None), empty)["mask"][0])
    seg_cases.append(dict(tile=None, n_obj=0, fail=None, keep=[], mask_bits=np.packbits(m_empty.astype(bool)), mask_max=int(m_empty.max()),
                          shape=tuple(m_empty.shape)))
    fx["seg_epilogue"] = seg_cases
    fx["collate"] = dict(inp=[{"a": torch.arange(3), "t": "x"}, {"a": torch.arange(2), "t": "y"}],
                         out=coll_ns["collate_fn_to_dict_list"]([{"a": torch.arange(3), "t": "x"}, {"a": torch.arange(2), "t": "y"}]))

    # ---- bucket wire format: real sender + real receiver ----
    d = extract("roll/utils/send_recv_utils.py", ["get_tensor_size", "TensorBucket", "SendBucketManager", "RecvBucketManager"])
    sr_ns = run_defs(d, base_ns())
    named = [("model.layers.0.mlp.down_proj.weight", torch.randn(96, 130, generator=g).bfloat16()), ("model.norm.weight", torch.randn(257, generator=g)),
             ("visual.blocks.0.attn.qkv.weight", torch.randn(60, 64, generator=g).bfloat16()), ("cnt", torch.arange(5, dtype=torch.int64))]
    tb_cuda_free = sr_ns["TensorBucket"].__init__.__defaults__  # (device="cuda",) -> construct on CPU explicitly
    mgr = sr_ns["SendBucketManager"].__new__(sr_ns["SendBucketManager"])
    mgr.bucket_size = 10_000
    mgr.bucket = sr_ns["TensorBucket"](10_000, device="cpu")
    buckets = []

    def emit(meta, buf):
        meta = {k: dict(v) for k, v in meta.items()}
        sr_ns["SendBucketManager"].meta_to_dict(meta)
        for m in meta.values():
            m["tensor_meta"] = {"shape": m["tensor_meta"]["shape"], "dtype": m["tensor_meta"]["dtype"]}
        buckets.append((meta, buf.clone()))

    for name, t in named:
        for meta, buf in mgr.push_tensor(t, name):
            emit(meta, buf)
    meta, buf = mgr.pop_last_bucket()
    if meta is not None:
        emit(meta, buf)
    rx = sr_ns["RecvBucketManager"]()
    got = {}
    for meta, buf in buckets:
        m2 = {k: dict(v, tensor_meta=torch.empty(v["tensor_meta"]["shape"], dtype=v["tensor_meta"]["dtype"], device="meta")) for k, v in meta.items()}
        got.update(rx.process_bucket(m2, buf))
    assert all(torch.equal(got[n_], t) for n_, t in named)   # the reference's own round trip
    fx["buckets"] = dict(named=named, bucket_size=10_000, buckets=buckets)

    dst = os.path.join(ROOT, "tests", "golden", "reference_intree.pt")
    torch.save(fx, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes;", {k: (len(v) if isinstance(v, list) else sorted(v)[:4]) for k, v in fx.items()})
    _ = tb_cuda_free, host_ref


if __name__ == "__main__":
    main()
