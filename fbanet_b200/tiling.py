"""Full-size tiled inference (SURVEY 8f-1 / BASELINE config 4): the reference's
``test_in_any_resolution.py:62-101`` driver with ``utils/dataset_utils.py`` ``tensor_divide_burst`` /
``tensor_merge`` as GPU gather / stitch kernels.

A burst ``[1,T,C,H,W]`` is reflect-padded to a multiple of ``psize`` (80), split into ``psize + 2*overlap``
(160) tiles with a reflected halo, each tile runs through the model, and the centre ``4*psize`` square of
every x4 output is stitched.  Tiles are independent units (the reference loops over them one at a time), so
they are sharded across ranks exactly like bursts; no activation halo is exchanged."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .dist import shard_range


@torch.no_grad()
def infer_full_resolution(model, burst: torch.Tensor, psize: int = 80, overlap: int = 40, tile_batch: int = 32,
                          rank: int = 0, world: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``burst [1,T,C,H,W]`` (CUDA fp32) -> ``[1,C,4H,4W]``.  With ``world > 1`` each rank fills only the
    rows/columns of its own tile shard (the rest of ``out`` is left untouched / zero)."""
    assert burst.dim() == 5 and burst.shape[0] == 1, "B=1 like the reference driver"
    _, T, C, H, W = burst.shape
    assert model.img_size == psize + 2 * overlap, "model must be built for the tile size"
    nh, nw = -(-H // psize), -(-W // psize)
    t0, t1 = shard_range(nh * nw, rank, world)
    if out is None:
        out = torch.zeros((C, 4 * H, 4 * W), device=burst.device, dtype=torch.float32)
    src = burst[0].contiguous().float()
    for b in range(t0, t1, tile_batch):
        e = min(b + tile_batch, t1)
        tiles = ops.tile_divide(src, psize, overlap, b, e)          # [n,T,C,160,160]
        sr = model(tiles)                                            # [n,C,640,640]
        ops.tile_merge(sr.contiguous(), out, H, W, psize, overlap, 4, b, e)
    return out.unsqueeze(0)


class BandedSession:
    """The symmetric-memory row bands of one full-size problem (input burst bands + x4 output bands), allocated and exchanged
    once and re-used for every image of that size: allocation + rendezvous cost ~100 ms, a multiple of the forward itself."""

    def __init__(self, T: int, C: int, H: int, W: int, device, group=None, scale: int = 4):
        import torch.distributed as dist

        from .dist import SymmetricBands, band_rows

        self.T, self.C, self.H, self.W, self.scale, self.group = T, C, H, W, scale, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.row0 = band_rows(H, self.world)
        self.src = SymmetricBands(T * C, self.row0, W, device, group)
        self.dst = SymmetricBands(C, [scale * r for r in self.row0], scale * W, device, group)


@torch.no_grad()
def infer_full_resolution_banded(model, band: torch.Tensor, H: int, W: int, psize: int = 80, overlap: int = 40,
                                 tile_batch: int = 32, group=None, gather_to: Optional[int] = 0,
                                 session: Optional[BandedSession] = None):
    """Config 4 with the burst SHARDED BY ROWS over the GPUs of one box (``BASELINE.json`` configs[3]): rank ``r`` holds
    ``band [T,C,rows_r,W]`` = image rows ``dist.band_rows(H, world)[r : r+2]`` and never the whole burst.

    The bands go into symmetric memory; each rank gathers ITS tiles (same contiguous tile shard as
    :func:`infer_full_resolution`) with one kernel that reads the rows of other bands -- the 40-pixel halo, or whole tile rows
    when the tile grid and the bands do not line up -- straight from the peers over NVLink, runs the model, and stores every
    x4 tile centre into the output band that owns those rows, again by peer stores.  No NCCL call on the data path.
    Returns this rank's output band ``[C, 4*rows_r, 4*W]``; with ``gather_to = g`` rank ``g`` instead returns the whole
    ``[1,C,4H,4W]`` image, copied band by band from peer memory (other ranks return their band).  Pass a
    :class:`BandedSession` to re-use the symmetric buffers across images of one size."""
    T, C, rows, Wb = band.shape
    dev = band.device
    ses = session if session is not None else BandedSession(T, C, H, W, dev, group)
    assert (ses.T, ses.C, ses.H, ses.W) == (T, C, H, W)
    world, rank, row0, src, dst = ses.world, ses.rank, ses.row0, ses.src, ses.dst
    assert Wb == W and rows == row0[rank + 1] - row0[rank], f"band {tuple(band.shape)} is not rows {row0[rank]}:{row0[rank + 1]} of {H}x{W}"
    assert model.img_size == psize + 2 * overlap, "model must be built for the tile size"
    src.barrier()                                                    # nobody still reads the previous image's bands
    src.local.copy_(band.reshape(T * C, rows, W))
    src.barrier()                                                    # every band is in place before anyone gathers a halo
    nh, nw = -(-H // psize), -(-W // psize)
    t0, t1 = shard_range(nh * nw, rank, world)
    for b in range(t0, t1, tile_batch):
        e = min(b + tile_batch, t1)
        tiles = ops.tile_divide_banded(src.ptrs, row0, T, C, H, W, psize, overlap, b, e, dev)     # gather + halo exchange
        sr = model(tiles)
        ops.tile_merge_banded(sr.contiguous(), dst.ptrs, row0, H, W, psize, overlap, 4, b, e)      # stitch into the owners' bands
    dst.barrier()                                                    # all peers' stores into my band have landed
    if gather_to is not None and rank == gather_to:
        out = torch.empty((C, 4 * H, 4 * W), device=dev, dtype=torch.float32)
        for k in range(world):
            out[:, 4 * row0[k]: 4 * row0[k + 1]].copy_(dst.peer(k))
        dst.barrier()                                                # peers keep their bands unchanged until the copy is done
        torch.cuda.current_stream(dev).synchronize()
        return out.unsqueeze(0)
    if gather_to is not None:
        dst.barrier()
    res = dst.local.clone()
    torch.cuda.current_stream(dev).synchronize()                     # nothing of this rank still reads peer memory when the bands are freed
    return res
