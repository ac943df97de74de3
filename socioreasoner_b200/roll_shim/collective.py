"""Named collective groups for the train -> infer weight broadcast (the reference's ``roll/utils/collective/collective.py``
as used by ``worker_helper.py:64-102``): a group is created from (world_size, rank, master_addr, master_port) independently
of the default ``torch.distributed`` world -- the sender (rank 0 = the train worker holding the weights) and the rollout
workers do not share a default process group.  NCCL over NVLink on the GPUs; gloo on CPU (tests).

Wire compatibility: the UNMODIFIED reference sender builds its side with ``init_custom_process_group``
(``roll/utils/collective/pg_utils.py:10-77``): a ``tcp://`` rendezvous, then ``PrefixStore(group_name, store)``, then torch's
``_new_process_group_helper(..., group_name=group_name)``.  The keys under which the ncclUniqueId / gloo addresses are exchanged
are derived from that prefix and that group name, so this side goes through the very same torch entry points with the same
arguments (``tests/test_weight_sync.py::test_group_rendezvous_with_reference_sender`` pairs the two)."""
from __future__ import annotations

from datetime import timedelta
from typing import Dict

import torch
import torch.distributed as dist

_groups: Dict[str, object] = {}


def init_collective_group(world_size: int, rank: int, backend: str = "nccl", group_name: str = "default",
                          master_addr: str = "127.0.0.1", master_port: int = 29600, timeout_s: float = 600.0):
    from torch.distributed import distributed_c10d as c10d

    if not group_name:
        raise ValueError(f"group_name {group_name!r} needs to be a non-empty string")
    if group_name in _groups:
        raise RuntimeError(f"collective group {group_name!r} already exists")
    assert 0 <= rank < world_size
    timeout = timedelta(seconds=timeout_s)
    store, rank, world_size = next(c10d.rendezvous(f"tcp://{master_addr}:{int(master_port)}", rank, world_size, timeout=timeout))
    store.set_timeout(timeout)
    store = c10d.PrefixStore(group_name, store)
    opt_name = "backend_options" if tuple(int(v) for v in torch.__version__.split("+")[0].split(".")[:2]) >= (2, 6) else "pg_options"
    pg, _ = c10d._new_process_group_helper(world_size, rank, [], c10d.Backend(backend), store, group_name=group_name,
                                           timeout=timeout, **{opt_name: None})
    c10d._world.pg_group_ranks[pg] = {i: i for i in range(world_size)}  # group rank == "global" rank inside this stand-alone group
    _groups[group_name] = pg
    return pg


def is_group_exist(group_name: str) -> bool:
    return group_name in _groups


def destroy_collective_group(group_name: str) -> None:
    _groups.pop(group_name, None)


def broadcast(tensor: torch.Tensor, src_rank: int = 0, group_name: str = "default") -> None:
    dist.broadcast(tensor, src=src_rank, group=_groups[group_name])


def allreduce(tensor: torch.Tensor, group_name: str = "default") -> None:
    dist.all_reduce(tensor, group=_groups[group_name])
