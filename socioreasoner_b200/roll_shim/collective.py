"""Named collective groups for the train -> infer weight broadcast (the reference's ``roll/utils/collective/collective.py``
as used by ``worker_helper.py:64-102``): a group is created from (world_size, rank, master_addr, master_port) independently
of the default ``torch.distributed`` world -- the sender (rank 0 = the train worker holding the weights) and the rollout
workers do not share a default process group.  NCCL over NVLink on the GPUs; gloo on CPU (tests)."""
from __future__ import annotations

from datetime import timedelta
from typing import Dict

import torch
import torch.distributed as dist

_groups: Dict[str, object] = {}


def init_collective_group(world_size: int, rank: int, backend: str = "nccl", group_name: str = "default",
                          master_addr: str = "127.0.0.1", master_port: int = 29600, timeout_s: float = 600.0):
    """Rendezvous through a TCP store at master_addr:master_port (rank 0 hosts it) and build a stand-alone backend."""
    if group_name in _groups:
        raise RuntimeError(f"collective group {group_name!r} already exists")
    timeout = timedelta(seconds=timeout_s)
    store = dist.TCPStore(master_addr, int(master_port), world_size, is_master=(rank == 0), timeout=timeout, wait_for_workers=True)
    store = dist.PrefixStore(f"sr_group/{group_name}", store)
    if backend == "gloo":
        pg = dist.ProcessGroupGloo(store, rank, world_size, timeout=timeout)
    elif backend == "nccl":
        opts = dist.ProcessGroupNCCL.Options()
        opts._timeout = timeout
        pg = dist.ProcessGroupNCCL(store, rank, world_size, opts)
    else:
        raise ValueError(f"unsupported backend {backend!r}")
    _groups[group_name] = pg
    return pg


def is_group_exist(group_name: str) -> bool:
    return group_name in _groups


def destroy_collective_group(group_name: str) -> None:
    _groups.pop(group_name, None)


def broadcast(tensor: torch.Tensor, src_rank: int = 0, group_name: str = "default") -> None:
    opts = dist.BroadcastOptions()
    opts.rootRank = src_rank
    opts.rootTensor = 0
    _groups[group_name].broadcast([tensor], opts).wait()


def allreduce(tensor: torch.Tensor, group_name: str = "default") -> None:
    _groups[group_name].allreduce([tensor], dist.AllreduceOptions()).wait()
