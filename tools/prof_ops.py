#!/usr/bin/env python
"""Micro-benchmarks of the non-GEMM kernels at the cfg2 (batch 64) shapes: window attention, LeFF depthwise
conv, LayerNorm, FAF gate, head conv.  `--case NAME --reps 1` for ncu captures."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbanet_b200 import ops, _lib as L  # noqa: E402

BF = torch.bfloat16
dev = torch.device("cuda:0")


def timeit(name, fn, reps, gbytes):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms:8.3f} ms   alg {gbytes:6.2f} GB -> {gbytes / ms:7.2f} TB/s", flush=True)


def attn(B, S, C, heads, shift):
    qkv = (torch.rand(B * S * S, 3 * C, device=dev) - 0.5).to(BF)
    table = torch.rand(361, heads, device=dev) * 0.1
    bx = ops.expand_rel_pos_bias(table, 10)
    bw = ops.expand_rel_pos_bias_wrap(bx, 10) if shift else None
    return (lambda: ops.window_attention(qkv, table, B, S, S, heads, 10, shift, (C // heads) ** -0.5, bias_expanded=bx, q_prescaled=True, bias_wrap=bw)), 2 * B * S * S * 4 * C / 1e9


def dw(B, S, C):
    x = (torch.rand(B, S, S, C, device=dev) - 0.5).to(BF)
    w, b = torch.rand(9, C, device=dev) - 0.5, torch.rand(C, device=dev)
    return (lambda: ops.dwconv3x3(x, w, b, L.ACT_GELU_TANH)), 2 * 2 * x.numel() / 1e9


def ln(B, S, C):
    x = (torch.rand(B * S * S, C, device=dev) - 0.5).to(BF)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    return (lambda: ops.layernorm(x, g, b)), 2 * 2 * x.numel() / 1e9


def gate(B, S):
    feat = (torch.rand(B, 14, S, S, 64, device=dev) - 0.5).to(BF)
    ws = torch.rand(9, 64, device=dev) - 0.5
    return (lambda: ops.faf_gate(feat, ws, want_gate=True, want_gated=True)), 2 * 2 * feat.numel() / 1e9


def head(B, S):
    x = torch.rand(B * 14, 3, S, S, device=dev)
    w, b = torch.rand(27, 64, device=dev) - 0.5, torch.rand(64, device=dev)
    return (lambda: ops.head_conv(x, w, b, BF)), (x.numel() * 4 + B * 14 * S * S * 64 * 2) / 1e9


def head_warp(B, S):
    """K1 fused into K0: the head conv sampling the unregistered burst through the homographies."""
    x = torch.rand(B * 14, 3, S, S, device=dev)
    w, b = torch.rand(27, 64, device=dev) - 0.5, torch.rand(64, device=dev)
    M = torch.eye(3, dtype=torch.float64).repeat(B, 14, 1, 1)
    M[:, 1:, :2, 2] = torch.rand(B, 13, 2, dtype=torch.float64) * 8 - 4
    M[:, 1:, 2, :2] = (torch.rand(B, 13, 2, dtype=torch.float64) - 0.5) * 2e-5
    M = M.view(-1, 3, 3).to(dev)
    return (lambda: ops.head_conv(x, w, b, BF, M=M, frames_per_burst=14)), (x.numel() * 4 + B * 14 * S * S * 64 * 2) / 1e9


def assemble(B, S):
    """Final assembly: hi/lo-folded fp32 SR columns [B,4S,4S,4] + bilinear x4 of the planar base frame -> planar fp32 image."""
    sr = torch.rand(B, 4 * S, 4 * S, 4, device=dev)
    base = torch.rand(B, 3, S, S, device=dev)
    return (lambda: ops.assemble(sr, base, 3)), (sr.numel() + 3 * B * 16 * S * S) * 4 / 1e9


def warp(B, S, C=3):
    x = torch.rand(B, 14, C, S, S, device=dev)
    M = torch.eye(3, dtype=torch.float64).repeat(B, 14, 1, 1)
    M[:, 1:, :2, 2] = torch.rand(B, 13, 2, dtype=torch.float64) * 8 - 4
    M[:, 1:, 2, :2] = (torch.rand(B, 13, 2, dtype=torch.float64) - 0.5) * 2e-5
    M = M.to(dev)
    return (lambda: ops.warp_burst(x, M)), 2 * x.numel() * 4 / 1e9


def flow(B, S):
    x = torch.rand(B, 14, 3, S, S, device=dev)
    fl = (torch.rand(B, 13, S, S, 2, device=dev) * 8 - 4)
    return (lambda: ops.flow_warp_burst(x, fl)), (2 * x.numel() + fl.numel()) * 4 / 1e9


def leff(B, S, C, f16=False):
    Hd = 4 * C
    T = torch.float16 if f16 else BF
    h1 = (torch.rand(B, S, S, Hd, device=dev) - 0.5).to(T)
    dw, db = torch.rand(9, Hd, device=dev) - 0.5, torch.rand(Hd, device=dev)
    w2, b2 = ((torch.rand(C, Hd, device=dev) - 0.5) * 0.05).to(T), torch.zeros(C, device=dev)
    res = torch.zeros(B, S, S, C, device=dev, dtype=BF)
    out = torch.empty_like(res)
    return (lambda: ops.leff_fc2(h1, dw, db, w2, b2, out, res, L.ACT_GELU_TANH)), 2 * (h1.numel() + 2 * res.numel()) / 1e9


def mlp(B, S, C, f16=False):
    """The one-kernel LeFF MLP (ops.leff_mlp) at a stage shape; algorithmic bytes = x in, residual in, out.  f16: fp16 fc2 weights =
    fp16 hidden tile + half2 depthwise path."""
    Hd = 4 * C
    x = (torch.rand(B, S, S, C, device=dev) - 0.5).to(BF)
    w1, b1 = ((torch.rand(Hd, C, device=dev) - 0.5) * 0.1).to(BF), torch.rand(Hd, device=dev) * 0.1
    dwt, db = (torch.rand(9, Hd, device=dev) - 0.5) * 0.3, torch.rand(Hd, device=dev) * 0.1
    w2, b2 = ((torch.rand(C, Hd, device=dev) - 0.5) * 0.05).to(torch.float16 if f16 else BF), torch.zeros(C, device=dev)
    res = torch.zeros(B, S, S, C, device=dev, dtype=BF)
    out = torch.empty_like(res)
    return (lambda: ops.leff_mlp(x, w1, b1, dwt, db, w2, b2, out, res, L.ACT_GELU_TANH)), 2 * 3 * res.numel() / 1e9


def faf_fuse(B, S):
    """K2 in one pass (ops.faf_fuse): bytes = features read once + fused map written."""
    feat = (torch.rand(B, 14, S, S, 64, device=dev) - 0.5).to(BF)
    ws = ops.faf_fuse_score_weight((torch.rand(9, 64, device=dev) - 0.5) * 0.1)
    wf, bf_, al = ((torch.rand(64, 14 * 64, device=dev) - 0.5) * 0.05).to(BF), torch.zeros(64, device=dev), torch.full((1,), 0.1, device=dev)
    out = torch.empty(B, S, S, 64, device=dev, dtype=BF)
    return (lambda: ops.faf_fuse(feat, ws, wf, bf_, al, out, want_gate=True)), (feat.numel() + out.numel()) * 2 / 1e9


CASES = {
    "faf_fuse_160": lambda: faf_fuse(64, 160),
    "mlp16_dec1_128": lambda: mlp(64, 160, 128, True),
    "mlp16_enc1_128": lambda: mlp(64, 80, 128, True),
    "mlp16_enc0_64": lambda: mlp(64, 160, 64, True),
    "mlp_dec1_128": lambda: mlp(64, 160, 128),
    "mlp_enc1_128": lambda: mlp(64, 80, 128),
    "mlp_enc0_64": lambda: mlp(64, 160, 64),
    "leff_dec1_128": lambda: leff(64, 160, 128),
    "leff_dec0_256": lambda: leff(64, 80, 256),
    "leff16_dec0_256": lambda: leff(64, 80, 256, True),
    "leff_enc0_64": lambda: leff(64, 160, 64),
    "attn_dec1_128x8_s5": lambda: attn(64, 160, 128, 8, 5),
    "attn_dec1_128x8_s0": lambda: attn(64, 160, 128, 8, 0),
    "attn_dec0_256x16_s5": lambda: attn(64, 80, 256, 16, 5),
    "attn_enc0_64x1_s5": lambda: attn(64, 160, 64, 1, 5),
    "attn_enc0_64x1_s0": lambda: attn(64, 160, 64, 1, 0),
    "attn_enc1_128x2_s5": lambda: attn(64, 80, 128, 2, 5),
    "attn_bott_256x16_s5": lambda: attn(64, 40, 256, 16, 5),
    "dw_160_512": lambda: dw(64, 160, 512),
    "dw_160_256": lambda: dw(64, 160, 256),
    "dw_80_1024": lambda: dw(64, 80, 1024),
    "ln_160_128": lambda: ln(64, 160, 128),
    "ln_160_64": lambda: ln(64, 160, 64),
    "ln_80_256": lambda: ln(64, 80, 256),
    "gate_160": lambda: gate(64, 160),
    "head_160": lambda: head(64, 160),
    "head_warp_160": lambda: head_warp(64, 160),
    "warp_160": lambda: warp(64, 160),
    "warp_raw80": lambda: warp(64, 80, 4),
    "assemble_160": lambda: assemble(64, 160),
    "flow_160": lambda: flow(64, 160),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    for n in (CASES if a.case == "all" else a.case.split(",")):
        fn, gb = CASES[n]()
        timeit(n, fn, a.reps, gb)
        del fn
        torch.cuda.empty_cache()
