#!/usr/bin/env python
"""BASELINE config 3 measurement: RealBSR-RAW shape -- 14-frame 4-channel packed-Bayer 80x80 bursts, homography warp (K1) + the
BaseModel forward (FAF fusion + SR x4 -> 320x320), batch 64 on one B200, inputs resident in HBM, CUDA-graph replay.
Synthetic inputs per SURVEY 8d: burst = rand(64,14,4,80,80), H_f = I + eps (translation U(-4,4) px, affine U(-0.01,0.01),
perspective U(-1e-5,1e-5), seed 1), H_0 = I.  Also checks one burst against the CPU oracle (checker only).  One JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbanet_b200 import BaseModel, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = dict(num_frames=14, img_size=80, in_channels=4, embed_dim=64, window_length=10)
    B, T, C, S = a.batch, 14, 4, 80
    model = BaseModel(**cfg, token_projection="linear", token_mlp="leff", dtype=a.dtype, seed=0).to(dev).eval()
    x = torch.rand(B, T, C, S, S, generator=torch.Generator().manual_seed(0)).to(dev)
    g = torch.Generator().manual_seed(1)
    M = torch.eye(3, dtype=torch.float64).repeat(B, T, 1, 1)
    M[:, 1:, :2, 2] = torch.rand(B, T - 1, 2, generator=g, dtype=torch.float64) * 8 - 4
    M[:, 1:, :2, :2] += torch.rand(B, T - 1, 2, 2, generator=g, dtype=torch.float64) * 0.02 - 0.01
    M[:, 1:, 2, :2] = torch.rand(B, T - 1, 2, generator=g, dtype=torch.float64) * 2e-5 - 1e-5
    Md = M.to(dev)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        def step():
            return model(ops.warp_burst(x, Md))
        for _ in range(a.warmup):
            y = step()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            y = step()
        graph.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            graph.replay()
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        fam = ops.profile_ops(step, stream, by_tag=False)
    # parity of burst 0 against the CPU oracle (same weights), fp32 path for the tolerance of the north-star
    from oracle.fbanet_oracle import build_oracle, psnr, warp_burst as warp_ref
    o = build_oracle(0, **cfg)
    xw_ref = torch.from_numpy(np.asarray(warp_ref(x[0].permute(0, 2, 3, 1).cpu().numpy(), M[0].numpy()), np.float32)).permute(0, 3, 1, 2)[None].contiguous()
    with torch.no_grad():
        ref = o(xw_ref)
    m32 = BaseModel(**cfg, token_projection="linear", token_mlp="leff", dtype="fp32", seed=0).to(dev).eval()
    m32.load_state_dict(o.state_dict())
    model.load_state_dict(o.state_dict())
    y32 = m32(ops.warp_burst(x[:1], Md[:1])).cpu()
    y16 = model(ops.warp_burst(x[:1], Md[:1])).cpu()
    warp_ms = fam.get("fbanet_warp_sm100", (0.0, 0, 0))[0]
    res = {"config": "cfg3: RealBSR-RAW shape, 14x4x80x80 packed-Bayer bursts, homography warp + FAF fusion + SR x4 -> 320x320",
           "n_gpus": 1, "batch": B, "dtype": a.dtype, "ms_per_step": ms, "bursts_per_s": B / (ms / 1e3), "output_mp_per_s": B * 0.1024 / (ms / 1e3),
           "steps": a.steps, "warmup": a.warmup, "cuda_graph": True, "warp_ms": warp_ms,
           "warp_gbs": (2 * x.numel() * 4 / 1e9) / (warp_ms / 1e3) if warp_ms else None,
           "parity_burst0": {"fp32_max_abs_vs_oracle": float((y32 - ref).abs().max()), "bf16_psnr_vs_oracle_db": float(psnr(y16, ref)),
                             "warp_max_abs_vs_oracle": float((ops.warp_burst(x[:1], Md[:1]).cpu() - xw_ref).abs().max())}}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
