#!/usr/bin/env python
"""BASELINE config 4 on N GPUs of one box: a full-size 14 x 3 x H x W burst sharded by ROW BANDS, tiles gathered with the halo read
from peer memory over NVLink (fbanet_tile_divide_banded_sm100), x4 tile centres stored into the owners' output bands
(fbanet_tile_merge_banded_sm100), no collective on the data path.  Launch with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_fullres.py

Prints one JSON line on rank 0 (device-timed, max over ranks) and, with --check, compares the stitched image with the single-GPU
replicated-burst driver (`infer_full_resolution`) bit for bit."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbanet_b200 import BaseModel  # noqa: E402
from fbanet_b200.dist import band_rows, halo_sources, init_from_env, reduce_max, shard_range  # noqa: E402
from fbanet_b200.tiling import BandedSession, infer_full_resolution, infer_full_resolution_banded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=1080)
    ap.add_argument("--W", type=int, default=1920)
    ap.add_argument("--frames", type=int, default=14)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--tile-batch", type=int, default=64)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank, local, world = init_from_env("nccl")
    assert world > 1, "run under torchrun with >= 2 ranks (the single-GPU driver is infer_full_resolution)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    H, W, T, C = a.H, a.W, a.frames, 3
    model = BaseModel(num_frames=T, img_size=160, in_channels=C, embed_dim=64, window_length=10, token_projection="linear",
                      token_mlp="leff", dtype=a.dtype, seed=0).to(dev)
    row0 = band_rows(H, world)
    # synthetic burst, seeded: every rank draws the same image on the host but uploads ONLY its band
    full = torch.rand(T, C, H, W, generator=torch.Generator().manual_seed(0))
    band = full[:, :, row0[rank]:row0[rank + 1]].contiguous().to(dev)
    if not (a.check and rank == 0):
        del full

    session = BandedSession(T, C, H, W, dev)

    def step():
        return infer_full_resolution_banded(model, band, H, W, tile_batch=a.tile_batch, gather_to=0, session=session)

    out = None
    for _ in range(a.warmup):
        out = step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = reduce_max(e0.elapsed_time(e1) / a.steps, dev)
    nh, nw = -(-H // 80), -(-W // 80)
    t0, t1 = shard_range(nh * nw, rank, world)
    need = halo_sources(H, 80, 40, (t0 // nw, (t1 - 1) // nw + 1), row0)
    remote_rows = sum(v for k, v in need.items() if k != rank)
    remote = torch.tensor([remote_rows * W * T * C * 4], dtype=torch.float64, device=dev)
    dist.all_reduce(remote)
    res = None
    if rank == 0:
        res = {"config": f"cfg4: {T}x{C}x{H}x{W} burst x4 -> {4 * H}x{4 * W}, {nh * nw} tiles of 160x160 (psize 80, overlap 40), row-band sharded",
               "n_gpus": world, "dtype": a.dtype, "ms_per_image": ms, "output_mp_per_s": 16 * H * W / 1e6 / (ms / 1e3),
               "tiles_per_s": nh * nw / (ms / 1e3), "steps": a.steps, "warmup": a.warmup,
               "halo_bytes_read_from_peers_per_image": int(remote.item()),
               "data_path": "P2P loads/stores on symmetric memory inside the tile gather / stitch kernels; no NCCL collective"}
        if a.check:
            whole = full[None].to(dev)
            ref = infer_full_resolution(model, whole, tile_batch=a.tile_batch)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            ref = infer_full_resolution(model, whole, tile_batch=a.tile_batch)
            s1.record()
            torch.cuda.synchronize()
            res["single_gpu_replicated_ms_per_image"] = s0.elapsed_time(s1)
            res["max_abs_diff_vs_replicated_single_gpu"] = float((out - ref).abs().max().item())
            res["bit_identical"] = bool(torch.equal(out, ref))
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
