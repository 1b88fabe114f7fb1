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


def band_rows(H: int, world: int):
    """Row boundaries ``[r_0 = 0, ..., r_world = H]`` of the row bands a full-size burst is sharded into (config 4): band ``k``
    = image rows ``[r_k, r_k+1)`` lives on GPU ``k``.  Same contiguous-shard rule as :func:`shard_range`."""
    if world > H:
        raise ValueError(f"{world} bands for {H} rows")
    return [shard_range(H, k, world)[0] for k in range(world)] + [H]


def halo_sources(H: int, psize: int, overlap: int, tile_rows, row0):
    """Which bands a rank must read to build its tiles: for the tile rows ``[i0, i1)`` it owns, the set of source rows (after both
    reflections of ``utils/dataset_utils.py:5-58``) mapped to ``{band: number of rows read from it}``.  Host-side statement of
    what ``tile_divide_banded_kernel`` touches; used by the tests and to report the NVLink halo volume."""
    Hp = -(-H // psize) * psize

    def refl(i, n):
        i = -i if i < 0 else i
        return 2 * (n - 1) - i if i >= n else i

    rows = set()
    for i in range(*tile_rows):
        for ty in range(psize + 2 * overlap):
            rows.add(refl(refl(i * psize + ty - overlap, Hp), H))
    out = {}
    for y in rows:
        k = max(b for b in range(len(row0) - 1) if y >= row0[b])
        out[k] = out.get(k, 0) + 1
    return out


class SymmetricBands:
    """Row bands of an fp32 image stack in CUDA symmetric memory (``torch.distributed._symmetric_memory``): every rank allocates
    the same-sized buffer, the rendezvous maps all peers' buffers into this process, and ``ptrs[k]`` is rank k's band as a device
    address usable by OUR kernels -- loads and stores on it travel over NVLink / NVSwitch.  torch supplies the allocation and
    the handle exchange only; the data path is ``fbanet_tile_{divide,merge}_banded_sm100``."""

    def __init__(self, planes: int, row0, W: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        self.row0 = list(row0)
        self.world = len(row0) - 1
        self.rank = dist.get_rank(group)
        assert self.world == dist.get_world_size(group)
        self.planes, self.W = planes, W
        max_rows = max(row0[k + 1] - row0[k] for k in range(self.world))
        self.buf = symm.empty(planes * max_rows * W, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.rows = row0[self.rank + 1] - row0[self.rank]

    @property
    def local(self) -> torch.Tensor:
        """This rank's band as ``[planes, rows, W]``."""
        return self.buf[: self.planes * self.rows * self.W].view(self.planes, self.rows, self.W)

    def peer(self, k: int) -> torch.Tensor:
        rows = self.row0[k + 1] - self.row0[k]
        return self.handle.get_buffer(k, (self.planes, rows, self.W), torch.float32)

    def barrier(self):
        """All ranks' prior writes to any band are visible to every rank's later reads (device-side signal-pad barrier)."""
        self.handle.barrier()
