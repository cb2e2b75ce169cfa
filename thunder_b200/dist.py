"""Host-side plumbing of the multi-GPU path: one process per GPU, particles sharded contiguously per rank
(the reference's Database::split, src/Database.cpp:621-641), a single exchange step per iteration (the half-map
all-reduce, Reconstructor::allReduceF/T, src/Reconstructor.cpp:2350-2484).  torch.distributed is only the
bootstrap / barrier / timing transport; the data-path collective is NCCL inside libthunder_b200 (thb_allreduce).
Works with the gloo backend on CPU for the tests."""
from __future__ import annotations

import os


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous [start, end) of `n_items` for `rank`: the first n % world ranks get one extra item"""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def half_set_of(global_index):
    """golden-standard half sets: even particles -> A (slot 0), odd -> B (slot 1), fixed for the whole run"""
    return global_index % 2


def init(backend: str = "nccl", device_index: int | None = None):
    """torch.distributed from the torchrun environment; returns (rank, world)"""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl" and device_index is not None:
            torch.cuda.set_device(device_index)
            kw["device_id"] = torch.device("cuda", device_index)
        dist.init_process_group(backend, **kw)
    return rank, world


def share_unique_id(make_id, rank: int, world: int) -> bytes | None:
    """rank 0 creates the NCCL unique id (thb_comm_unique_id), every rank receives it"""
    if world == 1:
        return None
    import torch.distributed as dist
    ids = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]


def max_over_ranks(values, device="cpu"):
    """element-wise maximum of a list of floats over all ranks (timings are the max over ranks)"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(values, device="cpu"):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]
