"""Session placement for multi-GPU serving: replicas only.

Dialogue sessions are independent (every state tensor is ``[B, ...]`` and every reduction is per
row), so N GPUs run N replicas of the weights, each owning a contiguous shard of the session slots;
there is no data-path collective (the reference deploys one process per GPU behind a
load balancer, ``swarm-config.yml:48-63``).  ``torch.distributed`` is used only to launch one
process per GPU and to reduce timings / counts in the benchmark.
"""
from __future__ import annotations

import os

import torch


def shard_sessions(total: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced shard of session slots [0, total) for ``rank``."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def init_distributed(backend: str | None = None) -> tuple[int, int]:
    """(rank, world_size) from the torchrun environment; initialises the process group if world > 1."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return rank, world


def _reduce(value: float, op) -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value: float) -> float:
    import torch.distributed as dist
    return _reduce(value, dist.ReduceOp.MAX)


def sum_over_ranks(value: float) -> float:
    import torch.distributed as dist
    return _reduce(value, dist.ReduceOp.SUM)


def barrier() -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
