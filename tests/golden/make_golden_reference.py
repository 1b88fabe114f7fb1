#!/usr/bin/env python
"""Golden vectors produced by EXECUTING the reference's own code (only possible in the build container, where the reference is
mounted at /root/reference): the torch-only `fba_net/utils/dataset_utils.py` (`tensor_divide_burst`, `tensor_merge` -- the
full-size tiling of test_in_any_resolution.py:62-101) is loaded by file path and run on a seeded burst whose size needs both
the reflect padding to a multiple of psize and the reflected halo.  The JAX model itself cannot be executed, so
this pins SURVEY 8f-1 only.  Run from the repo root: ``python tests/golden/make_golden_reference.py``."""
import contextlib
import importlib.util
import io
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/fba_net/utils/dataset_utils.py"

spec = importlib.util.spec_from_file_location("ref_dataset_utils", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

T, C, H, W, ps, ov, scale = 2, 2, 26, 38, 12, 6, 4
g = torch.Generator().manual_seed(123)
burst = torch.rand(1, T, C, H, W, generator=g)
with contextlib.redirect_stdout(io.StringIO()):       # the reference prints the padded shape
    blocks = ref.tensor_divide_burst(burst, ps, ov)
tiles = torch.stack([b[0] for b in blocks])           # [tiles, T, C, ps+2ov, ps+2ov]


def synthetic_sr(n, C, side):
    """Deterministic x4 'SR tiles' (an arithmetic pattern, so the fixture need not store them)."""
    k, c, y, x = np.meshgrid(np.arange(n), np.arange(C), np.arange(side), np.arange(side), indexing="ij")
    return torch.from_numpy((((k * 7 + c * 3 + y * 5 + x * 11) % 97) / 97.0).astype(np.float32))


sr = synthetic_sr(tiles.shape[0], C, scale * (ps + 2 * ov))
merged = ref.tensor_merge([t for t in sr], torch.zeros(1, C, scale * H, scale * W), scale * ps, scale * ov)
np.savez_compressed(os.path.join(HERE, "tiling_reference.npz"), burst=burst.numpy(), tiles=tiles.numpy(), merged=merged.numpy(),
                    psize=ps, overlap=ov, scale=scale)
print("reference tiling fixtures:", tuple(tiles.shape), tuple(merged.shape))

# ---- training loss (8f-3): the reference's own losses.py (torch-only), value + autograd gradient on a seeded pair whose values
# leave [0, 1] (clamp branches) and contain exact ties (sign(0) branches)
spec = importlib.util.spec_from_file_location("ref_losses", "/root/reference/fba_net/losses.py")
ref_losses = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_losses)
g = torch.Generator().manual_seed(321)
lx = (torch.rand(2, 3, 14, 18, generator=g) * 1.3 - 0.15)
ly = (torch.rand(2, 3, 14, 18, generator=g) * 1.3 - 0.15)
ly[0, 0, 3:6, 4:9] = lx[0, 0, 3:6, 4:9]                      # identical patch: zero differences and zero Sobel responses
lx.requires_grad_(True)
c1, c2 = ref_losses.CharbonnierLoss(), ref_losses.GWLoss(rgb_range=1.0)
lc, lg = c1(lx, ly), c2(lx, ly)
total = lc + 3 * lg                                          # train.py.bak:168
total.backward()
np.savez_compressed(os.path.join(HERE, "loss_reference.npz"), x=lx.detach().numpy(), y=ly.numpy(), charbonnier=lc.item(), gw=lg.item(),
                    total=total.item(), grad=lx.grad.numpy())
print("reference loss fixtures:", lc.item(), lg.item(), total.item())
