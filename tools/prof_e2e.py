#!/usr/bin/env python
"""Where does the end-to-end (host buffer) time go?  Raw pinned H2D / D2H rates on this box, alone and overlapped with the
forward, and a per-phase timeline of BaseModel.infer_host."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbanet_b200 import BaseModel  # noqa: E402

dev = torch.device("cuda:0")
CFG = dict(num_frames=14, img_size=160, in_channels=3, embed_dim=64, window_length=10)
B = 64
m = BaseModel(**CFG, token_projection="linear", token_mlp="leff", dtype="bf16", seed=0).to(dev).eval()
hin = torch.rand(B, 14, 3, 160, 160).pin_memory()
hout = torch.empty(B, 3, 640, 640).pin_memory()
xd = hin.to(dev)
yd = m(xd)
torch.cuda.synchronize()


def ev_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t = ev_time(lambda: xd.copy_(hin, non_blocking=True))
print(f"H2D {hin.numel() * 4 / 1e6:.0f} MB: {t:.2f} ms  {hin.numel() * 4 / t / 1e6:.1f} GB/s")
t = ev_time(lambda: hout.copy_(yd, non_blocking=True))
print(f"D2H {hout.numel() * 4 / 1e6:.0f} MB: {t:.2f} ms  {hout.numel() * 4 / t / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        xd.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2):
        hout.copy_(yd, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


print(f"H2D + D2H concurrently: {ev_time(both):.2f} ms")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    y2 = m(xd)
print(f"forward graph alone: {ev_time(g.replay):.2f} ms")
x2 = torch.empty_like(xd)


def fwd_and_copies():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        x2.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2):
        hout.copy_(yd, non_blocking=True)
    g.replay()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


print(f"forward graph with a full H2D and D2H running beside it: {ev_time(fwd_and_copies):.2f} ms")
for chunk in (32, (8, 24, 24, 8), (8, 28, 28), (16, 32, 16), (8, 16, 16, 16, 8), (4, 28, 28, 4), (16, 24, 24)):
    m.host_chunk = chunk
    m.infer_host(hin, hout)
    t0 = time.perf_counter()
    for _ in range(5):
        m.infer_host(hin, hout)
    print(f"infer_host chunk {chunk}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms wall")
