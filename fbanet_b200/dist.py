"""Multi-GPU plumbing: one process per GPU, bursts (or tiles) sharded across ranks, NO data-path collective.

The forward has no cross-burst term (``jax.vmap(model)``, reference ``train.py:35``), so ranks are
independent replicas over disjoint contiguous shards -- the same rule as the reference's DALI reader
(``pipeline/real_bsr_dataset.py:82-83``: ``shard_offset = shard_size * shard_id``).  ``torch.distributed``
is used only to rendezvous, to barrier around timed regions and to take the max of per-rank timings."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard ``[begin, end)`` of ``n_items`` for ``rank``; sizes differ by at most one and the
    shards tile ``range(n_items)`` exactly."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_from_env(backend: str = "nccl"):
    """(rank, local_rank, world) from torchrun's environment; initialises the process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def reduce_max(value: float, device=None) -> float:
    """Max over ranks of a per-rank scalar (device-timed milliseconds)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local: torch.Tensor, world_counts) -> torch.Tensor:
    """Concatenate per-rank row blocks on every rank (used only to assemble the cfg-4 image; outputs, not
    activations, cross ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    mx = max(world_counts)  # all_gather wants equal shapes: pad every block to the largest shard
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in world_counts]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, world_counts)], 0)
