#!/usr/bin/env python
"""BASELINE config 5 measurement: one data-parallel training step of BaseModel (2 bursts per GPU, 14x160x160 RGB ->
640x640 target, CharbonnierLoss + 3 GWLoss, AdamW(1e-4, wd 0.02)), `fbanet_b200.train.train_step` on the C-ABI kernels.
Single GPU:  python tools/run_train_step.py            N GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port P tools/run_train_step.py.  Timing: CUDA events around the K timed steps on the current
stream, barrier + synchronize on both sides, max over ranks; rank 0 prints one JSON line (bursts/s over all ranks, ms per step, the
loss before / after, peak memory, C-ABI launches per step).  NOT a bench.py line: the training step's kernels are first, CUDA-core
versions; this tool exists to profile them."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbanet_b200 import BaseModel, ops, train  # noqa: E402
from fbanet_b200 import dist as fdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2, help="bursts per GPU")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--img", type=int, default=160)
    ap.add_argument("--embed", type=int, default=64)
    ap.add_argument("--frames", type=int, default=14)
    ap.add_argument("--lr", type=float, default=1e-4)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = 0, 0
    if world > 1:
        rank, local, world = fdist.init_from_env("nccl")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    model = BaseModel(num_frames=a.frames, img_size=a.img, embed_dim=a.embed, window_length=10, token_projection="linear",
                      token_mlp="leff", dtype=a.dtype, seed=0).to(dev)
    for p in model.parameters():
        p.requires_grad_(True)
    flat = train.FlatParams(model.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    burst = torch.rand(a.batch, a.frames, 3, a.img, a.img, generator=g).to(dev)
    target = torch.rand(a.batch, 3, 4 * a.img, 4 * a.img, generator=g).to(dev)
    gen = torch.Generator().manual_seed(7 + rank)
    losses = []
    for _ in range(a.warmup):
        losses.append(train.train_step(model, flat, burst, target, a.lr, generator=gen)[0].item())
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = ops.LAUNCHES
    e0.record()
    for _ in range(a.steps):
        loss = train.train_step(model, flat, burst, target, a.lr, generator=gen)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "metric": "train_bursts_per_sec", "value": world * a.batch / (ms.item() * 1e-3), "unit": "bursts/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms.item(), "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": f"cfg5: train step, {a.batch} bursts/GPU {a.frames}x{a.img}x{a.img} RGB -> {4 * a.img}^2, embed {a.embed}",
                       "optimizer": "AdamW lr %g wd 0.02" % a.lr, "loss": "Charbonnier + 3 GW"},
            "loss_first": losses[0] if losses else None, "loss_last": loss[0].item(),
            "launches_per_step": (ops.LAUNCHES - before) // a.steps, "params": flat.numel,
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
