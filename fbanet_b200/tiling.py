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
