#!/usr/bin/env python
"""bf16-path error budget against the CPU oracle on the full cfg2 shape (one burst): per recorded stage the relative rms error
and the relative mean (systematic) error, for seeds / fold settings given on the command line.  Test infrastructure (imports oracle/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_model import FULL, _pair, _burst, _to_oracle_layout  # noqa: E402
from oracle.fbanet_oracle import psnr  # noqa: E402

dev = torch.device("cuda:0")
for seed in [int(a) for a in sys.argv[1:]] or [0]:
    o, m = _pair(FULL, "bf16", dev, seed=seed)
    x = _burst(FULL, 1, seed=seed)
    with torch.no_grad():
        ref = o.forward_stages(x)
    st = {}
    got = m.forward_stages(x.to(dev), st)
    gt = (torch.nn.functional.interpolate(x[:, 0], scale_factor=4, mode="bilinear", align_corners=False)
          + 0.05 * torch.randn(ref["out"].shape, generator=torch.Generator().manual_seed(7))).clamp(0, 1)
    g = got.cpu()
    print(f"seed {seed} fold_ln {m.fold_ln}: psnr delta {abs(psnr(g.clamp(0, 1), gt) - psnr(ref['out'].clamp(0, 1), gt)):.5f}  direct {psnr(g, ref['out']):.2f} dB"
          f"  mean err {(g - ref['out']).mean().item():+.2e}  rms err {(g - ref['out']).pow(2).mean().sqrt().item():.2e}")
    for k, v in st.items():
        if k not in ref or k == "faf.gate":
            continue
        r = ref[k]
        t = _to_oracle_layout(k, v, r)
        if t.shape != r.shape:
            continue
        e = t - r
        print(f"   {k:22s} rel rms {e.pow(2).mean().sqrt().item() / r.pow(2).mean().sqrt().item():.2e}   rel mean {e.mean().item() / r.abs().mean().item():+.2e}")
