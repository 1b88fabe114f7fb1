#!/usr/bin/env python
"""Micro-benchmark of the implicit-GEMM kernel on the layer shapes of the cfg2 forward (batch 64).
Used for ncu captures (`--case NAME --reps 1`) and quick CUDA-event timing of all cases."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbanet_b200 import ops, _lib as L  # noqa: E402

BF = torch.bfloat16
# name: (N, H, W, Cin, Cout, k, act, residual, store)
CASES = {
    "body3x3_64_64": (896, 160, 160, 64, 64, 3, L.ACT_RELU, False, L.STORE_NHWC),
    "body3x3_64_64_res": (896, 160, 160, 64, 64, 3, L.ACT_NONE, True, L.STORE_NHWC),
    "faf3x3_128_128": (64, 80, 80, 128, 128, 3, L.ACT_RELU, False, L.STORE_NHWC),
    "faf3x3_256_256": (64, 40, 40, 256, 256, 3, L.ACT_RELU, False, L.STORE_NHWC),
    "faf3x3_256_256_80": (64, 80, 80, 256, 256, 3, L.ACT_RELU, False, L.STORE_NHWC),
    "faf3x3_128_128_160": (64, 160, 160, 128, 128, 3, L.ACT_RELU, False, L.STORE_NHWC),
    "proj3x3_512_256": (64, 80, 80, 512, 256, 3, L.ACT_PRELU, False, L.STORE_NHWC),
    "tail3x3_64_256_320": (64, 320, 320, 64, 256, 3, L.ACT_NONE, False, L.STORE_CONVT2),
    "final3x3_64_16_640": (64, 640, 640, 64, 16, 3, L.ACT_NONE, False, L.STORE_NHWC),
    "fc1_128_512": (64, 160, 160, 128, 512, 1, L.ACT_GELU_TANH, False, L.STORE_NHWC),
    "fc2_512_128": (64, 160, 160, 512, 128, 1, L.ACT_NONE, True, L.STORE_NHWC),
    "qkv_128_384": (64, 160, 160, 128, 384, 1, L.ACT_NONE, False, L.STORE_NHWC),
    "fc1_256_1024": (64, 80, 80, 256, 1024, 1, L.ACT_GELU_TANH, False, L.STORE_NHWC),
    "fc1_256_1024_f16": (64, 80, 80, 256, 1024, 1, L.ACT_GELU_TANH, False, L.STORE_NHWC),      # fp16 store (the dim-256 LeFF's hidden map)
    "fc1_256_1024_40_f16": (64, 40, 40, 256, 1024, 1, L.ACT_GELU_TANH, False, L.STORE_NHWC),
    "qkv_256_768": (64, 80, 80, 256, 768, 1, L.ACT_NONE, False, L.STORE_NHWC),
    "qkv_64_192": (64, 160, 160, 64, 192, 1, L.ACT_NONE, False, L.STORE_NHWC),
    "proj_256_256": (64, 80, 80, 256, 256, 1, L.ACT_NONE, True, L.STORE_NHWC),
    "fc2_1024_256": (64, 80, 80, 1024, 256, 1, L.ACT_NONE, True, L.STORE_NHWC),
    "head_64_64": (896, 160, 160, 64, 64, 1, L.ACT_NONE, False, L.STORE_NHWC),
    "score3x3_64_2": (896, 160, 160, 64, 16, 3, L.ACT_NONE, False, L.STORE_NHWC_F32),     # FAF score conv: 2 fp32 outputs, tap-stacked
    "final3x3_64_3fold": (64, 640, 640, 64, 16, 3, L.ACT_NONE, False, -1),                  # final conv: hi/lo rows folded, tap-stacked
    "fuse_896_64": (64, 160, 160, 896, 64, 1, L.ACT_PRELU, False, L.STORE_NHWC),
}


def run_case(name, reps, dev):
    N, H, W, ci, co, k, act, res, store = CASES[name]
    g = torch.Generator(device=dev).manual_seed(0)
    x = (torch.rand(N, H, W, ci, device=dev, generator=g) - 0.5).to(BF)
    w = ((torch.rand(co, k * k * ci, device=dev, generator=g) - 0.5) * 0.05).to(BF)
    b = torch.zeros(co, device=dev)
    alpha = torch.full((1,), 0.25, device=dev)
    kw = {}
    if store == L.STORE_NHWC_F32:
        out = torch.empty(N, H, W, 2, device=dev, dtype=torch.float32)
        kw = dict(cout_store=2)
    elif store == -1:
        store = L.STORE_NHWC_F32
        out = torch.empty(N, H, W, 4, device=dev, dtype=torch.float32)
        kw = dict(cout_store=6, fold_hi_lo=True)
    elif store == L.STORE_CONVT2:
        out = torch.empty(N, 2 * H, 2 * W, co // 4, device=dev, dtype=BF)
    else:
        out = torch.empty(N, H, W, co, device=dev, dtype=torch.float16 if name.endswith("_f16") else BF)
    r = torch.zeros_like(out) if res else None

    def go():
        ops.conv_gemm([x], w, out, kh=k, kw=k, pad=k // 2, bias=b, act=act, alpha=alpha, residual=r, store_mode=store, impl=L.IMPL_TCGEN05, **kw)

    go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * N * H * W * k * k * ci * co
    byts = 2.0 * (x.numel() + out.numel() * (2 if res else 1))
    print(f"{name:22s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  min-HBM {byts / 1e9:6.2f} GB -> {byts / ms / 1e9 * 1e3 / 1e3:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    for n in (CASES if a.case == "all" else a.case.split(",")):
        run_case(n, a.reps, dev)
