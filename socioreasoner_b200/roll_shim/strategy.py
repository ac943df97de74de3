"""B200 strategies behind the reference's ``InferenceStrategy`` surface.

``B200InferStrategy`` answers to what ``HfInferStrategy`` (``roll/distributed/strategy/hf_strategy.py:26-129``) and the
rollout side of ``VllmStrategy`` (``vllm_strategy.py:114-141``) answer to; ``B200SegStrategy`` to ``SegInferStrategy``
(``seg_strategy.py:12-79``).  The micro-batch loop, the multi-modal input concatenation and the result collation follow
the reference line by line (cited inline) so that swapping the strategy does not change what the workers see.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from .protocol import DataProto, collate_fn_to_dict_list


class InferenceStrategy:
    """``roll/distributed/strategy/strategy.py:16-138`` (the members the hot path touches)."""
    strategy_name = None

    def __init__(self, worker):
        self.worker = worker
        self.model = None
        self.tokenizer = None
        self.worker_config = getattr(worker, "worker_config", None)
        self.model_update_comm_plan = {}

    def initialize(self, *args, **kwargs):
        raise NotImplementedError

    def forward_step(self, batch, forward_func):
        raise NotImplementedError

    def get_data_input(self, batch):
        return batch

    def generate(self, *args, **kwargs):
        raise NotImplementedError

    def unwrap_model(self):
        return self.model

    # offload / reload around every RPC (context_managers.py:126-128,175-177).  On a 180 GB B200 the 7.5 GB replica stays
    # resident: both are no-ops that keep the call contract.
    def load_states(self, include=None, non_blocking=False):
        return None

    def offload_states(self, include=None, non_blocking=False):
        return None

    # ---- weight sync, receive side (worker_helper.py:64-116 behind vllm_strategy.py:258-271) ----
    # The train worker (rank 0 of a named group) broadcasts every parameter under its HF checkpoint name, whole
    # (``broadcast_parameter``) or in int8 buckets (``broadcast_bucket``); co-located workers get the tensors directly
    # (``update_parameter*``).  Received tensors overwrite the kernel-layout weights in place (``model.load_weights``).
    def _worker_rank(self) -> int:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    def _sync_device(self):
        dev = getattr(self.model, "device", None)
        return torch.device(dev) if dev is not None else torch.device("cpu")

    def setup_collective_group(self, comm_plan, backend="nccl"):
        from . import collective
        rank_in_cluster = getattr(self.worker, "rank", 0)
        rank, args = get_dist_info_from_comm_plan(comm_plan, rank_in_cluster=rank_in_cluster, rank_in_worker=self._worker_rank())
        if rank is None:
            return
        world_size = len(args["tgt_devices"]) + 1
        collective.init_collective_group(world_size, rank, backend=backend, group_name=args["group_name"],
                                         master_addr=args["master_addr"], master_port=args["master_port"])
        warm = torch.zeros(1, device=self._sync_device())
        collective.allreduce(warm, group_name=args["group_name"])  # a small all_reduce for warm-up, as the reference does
        self.model_update_comm_plan[args["src_pp_rank"]] = dict(rank=rank, world_size=world_size, src_pp_rank=args["src_pp_rank"],
                                                                group_name=args["group_name"], comm_plan=comm_plan,
                                                                comm_plan_args=args)

    def broadcast_parameter(self, src_pp_rank, dtype, shape, parameter_name):
        from . import collective
        if src_pp_rank not in self.model_update_comm_plan:
            return
        weight = torch.empty(shape, dtype=dtype, device=self._sync_device())
        collective.broadcast(tensor=weight, src_rank=0, group_name=self.model_update_comm_plan[src_pp_rank]["group_name"])
        self.update_parameter(parameter_name, weight, [self._worker_rank()])

    def broadcast_bucket(self, src_pp_rank, meta_infos, bucket_size):
        from . import collective
        if src_pp_rank not in self.model_update_comm_plan:
            return
        buffer = torch.empty(bucket_size, dtype=torch.int8, device=self._sync_device())
        collective.broadcast(tensor=buffer, src_rank=0, group_name=self.model_update_comm_plan[src_pp_rank]["group_name"])
        self.update_parameter_in_bucket(meta_infos, buffer, [self._worker_rank()])

    def update_parameter(self, parameter_name, weight, ranks_in_worker):
        if self._worker_rank() not in ranks_in_worker:
            return
        self.model.load_weights([(parameter_name, weight)])

    def update_parameter_in_bucket(self, meta_infos, buffer, ranks_in_worker):
        from .send_recv import RecvBucketManager
        if self._worker_rank() not in ranks_in_worker:
            return
        if getattr(self, "recv_manager", None) is None:
            self.recv_manager = RecvBucketManager()
        named = self.recv_manager.process_bucket(meta_infos, buffer)
        self.model.load_weights(list(named.items()))

    # ---- distributed ops (strategy.py:123-138) on the fused device kernel ----
    def op_compute_log_probs(self, logits: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor):
        """logits [b,S,V], input_ids [b,S], attention_mask(response_mask) [b,S] -> [b,S-1] (strategy.py:123-133)."""
        from .. import lib
        b, S, V = logits.shape
        labels = input_ids[:, 1:].clone()
        labels[attention_mask[:, 1:] == 0] = 0  # avoid invalid token id
        lg = logits[:, :-1].reshape(-1, V)
        lp, _ = lib.logprob_entropy(lg, labels.reshape(-1).to(lg.device).contiguous(), want_entropy=False)
        return lp.view(b, S - 1) * attention_mask[:, 1:].to(lp.device)

    def op_compute_entropy(self, logits: torch.Tensor, attention_mask: torch.Tensor):
        from .. import lib
        b, S, V = logits.shape
        _, ent = lib.logprob_entropy(logits.reshape(-1, V), None, want_entropy=True)
        return ent.view(b, S)[:, :-1] * attention_mask[:, 1:].to(ent.device)


def get_dist_info_from_comm_plan(comm_plan, rank_in_cluster, rank_in_worker):
    """``roll/utils/functionals.py:875-882``: this worker's rank inside the broadcast group (sender = 0) and the plan entry."""
    for src_rank, comm_plan_args in comm_plan.items():
        start_rank = 0
        for tgt_device in comm_plan_args["tgt_devices"]:
            start_rank += 1
            if tgt_device["rank"] == rank_in_cluster and tgt_device["device"]["rank"] == rank_in_worker:
                return start_rank, comm_plan_args
    return None, None


def _concat_multi_modal(data: DataProto, forward_args: dict, device) -> dict:
    """hf_strategy.py:68-79: per-sample dicts -> concatenated pixel_values / image_grid_thw."""
    if "multi_modal_inputs" in data.non_tensor_batch:
        mm = defaultdict(list)
        for sample in data.non_tensor_batch["multi_modal_inputs"]:  # samples may be empty (text + mm mixed data)
            for key in sample.keys():
                mm[key].append(sample[key])
        for key in mm.keys():
            assert key not in forward_args
            forward_args[key] = torch.concat(mm[key], dim=0).to(device)
    return forward_args


class B200InferStrategy(InferenceStrategy):
    strategy_name = "b200_infer"
    fuse_log_probs = True  # forward_step(forward_func_log_probs) -> fused LM head + log-prob / entropy kernel

    def initialize(self, model_provider):
        """hf_strategy.py:34-47: sets ``self.tokenizer`` (``default_tokenizer_provider``, read back by the workers:
        base_worker.py:49,129-131), ``self.model`` and the worker's rank_info (dispatch depends on it, decorator.py:113-121).
        The processor (image min / max pixels) is loaded too: the vLLM-form rollout input carries PIL images, not pixel rows.
        One process per GPU; torch.distributed (NCCL) is initialised by the launcher."""
        import torch.distributed as dist
        from .model_providers import default_processor_provider, default_tokenizer_provider
        wc = self.worker_config
        model_args = getattr(wc, "model_args", None)
        if model_args is not None:
            self.tokenizer = default_tokenizer_provider(model_args=model_args)
            self.processor = default_processor_provider(model_args=model_args)
        self.model = model_provider(tokenizer=self.tokenizer, model_args=model_args, is_trainable=False)
        ri = getattr(self.worker, "rank_info", None)
        if ri is not None:
            ri.dp_rank = dist.get_rank() if dist.is_initialized() else 0
            ri.dp_size = dist.get_world_size() if dist.is_initialized() else 1
        return self

    def forward_step(self, batch: DataProto, forward_func: Callable[[DataProto, torch.Tensor], Tuple[torch.Tensor, Dict[str, torch.Tensor]]]
                     ) -> Dict[str, torch.Tensor]:
        """hf_strategy.py:49-94."""
        batch_size = batch.batch.batch_size[0]
        micro_batch_size = batch.meta_info["micro_batch_size"]
        num_microbatches = max(batch_size // micro_batch_size, 1)
        losses_reduced = []
        for data in batch.chunk(chunks=num_microbatches):
            input_ids = data.batch["input_ids"]
            attention_mask = data.batch["attention_mask"]
            position_ids = data.batch["position_ids"]
            forward_args = dict(data.meta_info.get("forward_args", {}))
            if position_ids.dim() == 3:
                position_ids = position_ids.transpose(0, 1)  # (bsz, 3, seqlen) -> (3, bsz, seqlen)
            forward_args = _concat_multi_modal(data, forward_args, input_ids.device)
            if self.fuse_log_probs and getattr(forward_func, "__name__", "") == "forward_func_log_probs" \
                    and hasattr(self.model, "compute_log_probs") and "response_mask" in data.batch.keys():
                # the stock forward_func (base_worker.py:236-246) only reduces the logits to log-probs and entropy under the
                # response mask: run it fused with the LM head, on the masked-in rows only (no [b, S, V] logits)
                fa = {k: v for k, v in forward_args.items() if k != "force_vit_image"}
                lp, ent = self.model.compute_log_probs(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                                       response_mask=data.batch["response_mask"], **fa)
                losses_reduced.append({"log_probs": lp, "entropy": ent})
                continue
            output = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids, use_cache=False,
                                **forward_args)
            loss, loss_reduced = forward_func(data, output.logits)
            losses_reduced.append(loss_reduced)
        return collate_fn_to_dict_list(losses_reduced)

    processor = None
    _preprocessor = None

    def _image_preprocessor(self):
        """Device image path (``sr_preprocess_tile``) with the pixel budget of the checkpoint's processor, the one the
        collator's token expansion was made with (collator.py:456-462); the reference-effective default otherwise."""
        if self._preprocessor is None:
            from ..preprocess import GpuPreprocessor, REFERENCE_MAX_PIXELS
            ip = getattr(self.processor, "image_processor", None)
            mn, mx = getattr(ip, "min_pixels", None), getattr(ip, "max_pixels", None)
            if mn is None or mx is None:  # transformers 5.x keeps them in `size`
                size = getattr(ip, "size", None) or {}
                mn, mx = size.get("shortest_edge", 56 * 56), size.get("longest_edge", REFERENCE_MAX_PIXELS)
            self._preprocessor = GpuPreprocessor(self._sync_device(), min_pixels=int(mn), max_pixels=int(mx))
        return self._preprocessor

    def _generate_engine_form(self, batch: DataProto, generation_config: dict) -> torch.Tensor:
        """The input form the pipeline actually sends (rlvr_socioseg_vlm_pipeline_infer.py:645-657, consumed by
        vllm_strategy.py:114-141): ``non_tensor_batch["multi_modal_data"][i] = {"prompt_token_ids": [...], "multi_modal_data":
        {"image": [PIL, ...]}}`` next to the left-padded ``input_ids`` / ``attention_mask`` / ``position_ids`` (already expanded
        to one placeholder per merged patch by the collator's processor call).  vLLM re-runs the image processor on the PIL
        images; here they go through the device preprocessing kernel.  Returns what ``VllmStrategy.generate`` returns:
        ``[b * n, S_in + max response length]``, responses right-padded with ``pad_token_id`` (stop token included)."""
        from torch.nn.utils.rnn import pad_sequence
        gc = generation_config
        if gc.get("num_beams", 1) > 1:
            raise NotImplementedError("beam search is not on the SocioSeg path")
        n = int(gc.get("num_return_sequences", 1))
        temperature = float(gc.get("temperature", 1.0))
        top_k = gc.get("top_k", 0)
        pad = gc["pad_token_id"]
        batch_size = batch.batch.batch_size[0]
        micro = batch.meta_info.get("micro_batch_size", batch_size)
        outs: List[torch.Tensor] = []
        dev = self._sync_device()
        for data in batch.chunk(chunks=max(batch_size // micro, 1)):
            images = []
            for item in data.non_tensor_batch["multi_modal_data"]:
                mm = item.get("multi_modal_data") if isinstance(item, dict) else None
                for img in (mm or {}).get("image", []):
                    images.append(np.asarray(img.convert("RGB") if hasattr(img, "convert") else img, dtype=np.uint8))
            pv, thw = (None, None)
            if images:
                pv, thw = self._image_preprocessor()(images, out_dtype=torch.bfloat16)
            pos = data.batch["position_ids"] if "position_ids" in data.batch else None
            if pos is not None and pos.dim() == 3:
                pos = pos.transpose(0, 1)  # DataProto keeps (bsz, 3, seqlen); the model takes (3, bsz, seqlen) (hf_strategy.py:65-67)
            seq = self.model.generate(input_ids=data.batch["input_ids"].to(dev), attention_mask=data.batch["attention_mask"].to(dev),
                                      position_ids=pos, use_cache=True, pixel_values=pv, image_grid_thw=thw,
                                      max_new_tokens=int(gc["max_new_tokens"]), do_sample=temperature > 0 and top_k != 1,
                                      temperature=temperature if temperature > 0 else 1.0, top_p=float(gc.get("top_p", 1.0)),
                                      top_k=0 if top_k in (None, -1) else int(top_k), num_return_sequences=n,
                                      repetition_penalty=float(gc.get("repetition_penalty", 1.0)),
                                      eos_token_id=gc.get("eos_token_id"), pad_token_id=pad)
            s_in = data.batch["input_ids"].shape[1]
            for row in seq[:, s_in:]:
                keep = int((row != pad).nonzero().max().item()) + 1 if bool((row != pad).any()) else 0
                outs.append(row[:keep])
        output_ids = pad_sequence(outs, batch_first=True, padding_value=pad)           # gather_outputs_to_pad_tensor
        input_ids = batch.batch["input_ids"].to(output_ids.device)
        b, s_in = input_ids.shape
        repeated = input_ids.unsqueeze(1).repeat(1, n, 1).view(b * n, s_in)           # concatenate_input_and_output
        return torch.cat((repeated, output_ids), dim=1)

    def generate(self, batch: DataProto, generation_config: dict) -> torch.Tensor:
        """hf_strategy.py:96-129: micro-batched ``model.generate`` then pad_sequence with pad_token_id; the vLLM input form
        (``multi_modal_data`` in the non-tensor batch) goes through ``_generate_engine_form``."""
        from torch.nn.utils.rnn import pad_sequence
        if "multi_modal_data" in batch.non_tensor_batch:
            return self._generate_engine_form(batch, generation_config)
        batch_size = batch.batch.batch_size[0]
        micro_batch_size = batch.meta_info["micro_batch_size"]
        num_microbatches = max(batch_size // micro_batch_size, 1)
        output_list: List[torch.Tensor] = []
        for data in batch.chunk(chunks=num_microbatches):
            forward_args = dict(data.meta_info.get("forward_args", {}))
            forward_args = _concat_multi_modal(data, forward_args, data.batch["input_ids"].device)
            output = self.model.generate(input_ids=data.batch["input_ids"], attention_mask=data.batch["attention_mask"],
                                         use_cache=True, **forward_args, **generation_config)
            output_list.extend(list(output))
        return pad_sequence(output_list, batch_first=True, padding_value=generation_config["pad_token_id"])


class B200SegStrategy(InferenceStrategy):
    """Mask side (seg_strategy.py:12-79).  ``initialize(model_provider)`` receives what the reference's provider returns
    (``sam2_seg_model_provider``, model_providers.py:515-562): an object with the ``SAM2ImagePredictor`` API --
    ``set_image(PIL 756x756)`` once per sample, ``predict(box=, point_coords=, point_labels=) -> (masks [3,H,W], scores [3],
    low_res_logits [3,256,256])`` per object.  This strategy owns everything the reference then does on the CPU: best-score
    candidate pick, the predictor's own post-processing of the picked candidate (bilinear up-sample of the low-res logits to the
    image size, > 0), OR over objects, cv2 nearest resize to 768, uint8 -- one fused kernel per sample (``sr_mask_epilogue``).
    Predictors are used in this order: the B200-native SAM2 (``socioreasoner_b200.sam2``: logits stay on the device), any object
    with the reference API (its third return value feeds the kernel; if it is None the thresholded masks do), or a plain callable
    ``predictor(image, prompt) -> (low_res_logits [3,h,w], scores [3])`` (synthetic logits in tests / bench)."""
    strategy_name = "seg_infer"

    def __init__(self, worker, image_size: int = 756, out_size: int = 768):
        super().__init__(worker)
        self.image_size, self.out_size = image_size, out_size
        self.predictor = None

    def initialize(self, model_provider):
        self.predictor = model_provider(model_args=getattr(self.worker_config, "model_args", None), is_trainable=False)
        self.model = self.predictor
        return self

    def _has_predictor_api(self) -> bool:
        return hasattr(self.predictor, "set_image") and hasattr(self.predictor, "predict")

    def _predict_one(self, image, single_vp):
        """-> (candidate logits [3, h, w] fp32, scores [3] fp32) for one object prompt."""
        if not self._has_predictor_api():
            l, s = self.predictor(image, single_vp)
            return torch.as_tensor(l, dtype=torch.float32), torch.as_tensor(s, dtype=torch.float32)
        prompt = {}
        if "point_coords" in single_vp and "point_labels" in single_vp:  # seg_strategy.py:51-56
            prompt["point_coords"] = single_vp["point_coords"]
            prompt["point_labels"] = single_vp["point_labels"]
        if "box" in single_vp:
            prompt["box"] = single_vp["box"]
        if hasattr(self.predictor, "predict_logits"):   # B200-native SAM2: candidate logits and scores stay on the device
            return self.predictor.predict_logits(**prompt)
        pred_masks, scores, low_res = self.predictor.predict(**prompt)
        scores = torch.as_tensor(np.asarray(scores) if not torch.is_tensor(scores) else scores, dtype=torch.float32)
        if low_res is not None:
            low = low_res if torch.is_tensor(low_res) else torch.from_numpy(np.asarray(low_res))
            return low.to(torch.float32), scores
        m = pred_masks if torch.is_tensor(pred_masks) else torch.from_numpy(np.asarray(pred_masks))
        return (m > 0).to(torch.float32) * 2.0 - 1.0, scores  # already thresholded at image size: +-1 "logits", identity up-sample

    def segment(self, batch: DataProto) -> dict:
        from .. import lib
        masks = []
        for i in range(len(batch)):
            image = batch.non_tensor_batch["seg_image"][i]
            visual_prompt = batch.non_tensor_batch["visual_prompt"][i]
            if len(visual_prompt) == 0:  # seg_strategy.py:40-45: zero mask, model untouched
                masks.append({"mask": np.zeros((self.out_size, self.out_size), dtype=np.uint8)})
                continue
            if self._has_predictor_api():
                if hasattr(image, "resize"):
                    image = image.resize((self.image_size, self.image_size))     # seg_strategy.py:37
                self.predictor.set_image(image)                                   # seg_strategy.py:47
            logits, scores = [], []
            for single_vp in visual_prompt:
                try:
                    l, s = self._predict_one(image, single_vp)
                    logits.append(l)
                    scores.append(s)
                except Exception:  # seg_strategy.py:62-63: a failing object is skipped
                    continue
            if not logits:
                # every object failed: the reference still resizes its zero 756 mask -> zero 768 mask
                masks.append({"mask": np.zeros((self.out_size, self.out_size), dtype=np.uint8)})
                continue
            m = lib.mask_epilogue(torch.stack(logits).cuda().contiguous(), torch.stack(scores).cuda().contiguous(), self.image_size,
                                  self.out_size)
            masks.append({"mask": m.cpu().numpy()})
        return collate_fn_to_dict_list(masks)


def create_strategy(worker):
    """``roll/distributed/strategy/factory.py:7-30`` with the B200 backends answering to the existing names, so the
    entry script and the yaml stay unchanged: rollout (`vllm`, `hf_infer`) and mask (`seg_infer`)."""
    name = worker.worker_config.strategy_args.strategy_name
    if name in ("b200_infer", "hf_infer", "vllm"):
        return B200InferStrategy(worker)
    if name in ("b200_seg", "seg_infer"):
        return B200SegStrategy(worker)
    raise ValueError(f"Unknown strategy name: {name}")
