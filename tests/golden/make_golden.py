#!/usr/bin/env python
"""Generates the committed fixtures from the CPU oracle (the reference has no golden vectors and cannot be
executed).  Run from the repo root: ``python tests/golden/make_golden.py``."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fbanet_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

cfg = dict(num_frames=4, img_size=40, in_channels=3, embed_dim=32, window_length=10)
m = O.build_oracle(0, **cfg)
x = torch.rand(1, 4, 3, 40, 40, generator=torch.Generator().manual_seed(0))
with torch.no_grad():
    st = m.forward_stages(x)
stats = {k: (v.mean().item(), v.std().item()) for k, v in st.items() if k != "out"}
torch.save({"seed": 0, "cfg": cfg, "x": x, "out": st["out"], "stage_stats": stats}, os.path.join(HERE, "small_model.pt"))

rng = np.random.default_rng(7)
burst = rng.random((4, 24, 32, 3)).astype(np.float32)
M = np.tile(np.eye(3), (4, 1, 1))
M[:, :2, :2] += rng.uniform(-0.01, 0.01, (4, 2, 2))
M[:, :2, 2] += rng.uniform(-4, 4, (4, 2))
M[:, 2, :2] += rng.uniform(-1e-5, 1e-5, (4, 2))
M[0] = np.eye(3)
np.savez_compressed(os.path.join(HERE, "warp.npz"), burst=burst, M=M, out=O.warp_burst(burst, M))
print("golden fixtures written to", HERE)
