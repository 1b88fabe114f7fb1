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
# the same pair through the trainer's two lines as written (train.py.bak:167-168): clamp(restored, 0, 1), then both criteria
lx2 = lx.detach().clone().requires_grad_(True)
restored = torch.clamp(lx2, 0, 1)
tc = c1(restored, ly) + 3 * c2(restored, ly)
tc.backward()
np.savez_compressed(os.path.join(HERE, "loss_reference.npz"), x=lx.detach().numpy(), y=ly.numpy(), charbonnier=lc.item(), gw=lg.item(),
                    total=total.item(), grad=lx.grad.numpy(), total_clamped=tc.item(), grad_clamped=lx2.grad.numpy())
print("reference loss fixtures:", lc.item(), lg.item(), total.item())

# ---- learning-rate schedules of the training configuration (8f-3): the reference's own `warmup_scheduler/scheduler.py` (torch-only)
# driven exactly as train.py.bak:104-115,220 drives it -- one step before the first epoch, one after every epoch -- recording the
# optimizer's learning rate in effect DURING each epoch
import warnings

spec = importlib.util.spec_from_file_location("ref_scheduler", "/root/reference/fba_net/warmup_scheduler/scheduler.py")
ref_sched = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_sched)
sched_out = {}
for tag, nepoch, warm, lr0 in (("warmup_default", 250, 3, 1e-4), ("warmup_short", 20, 5, 2e-4), ("steplr", 250, 0, 1e-4)):
    opt_ = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=lr0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if warm:
            cosine = torch.optim.lr_scheduler.CosineAnnealingLR(opt_, nepoch - warm, eta_min=1e-6)
            sched = ref_sched.GradualWarmupScheduler(opt_, multiplier=1, total_epoch=warm, after_scheduler=cosine)
        else:
            sched = torch.optim.lr_scheduler.StepLR(opt_, step_size=50, gamma=0.5)
        sched.step()
        lrs = []
        for epoch in range(1, nepoch + 1):
            lrs.append(opt_.param_groups[0]["lr"])
            opt_.step()
            sched.step()
    sched_out[tag] = np.array(lrs, dtype=np.float64)
    sched_out[tag + "_cfg"] = np.array([nepoch, warm, lr0], dtype=np.float64)
np.savez_compressed(os.path.join(HERE, "lr_schedule_reference.npz"), **sched_out)
print("reference lr schedules:", {k: (v[:6].round(8).tolist(), v[-1]) for k, v in sched_out.items() if not k.endswith("_cfg")})
