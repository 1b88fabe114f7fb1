#!/usr/bin/env python
"""Dry run of tests/test_zz_gpu_train_compositions.py WITHOUT a GPU: the test bodies are executed on the CPU with fbanet_b200.ops
replaced by the torch stand-ins of tests/test_host_logic.py (fp32 arithmetic; for the bf16 LeWin case stand-ins that compute in fp32
and round every op output to bf16, as the kernels do).  It checks the TEST code (shapes, reference expressions, tolerances with their
margins), not the kernels: written because those tests were added after round 1's GPU budget was spent.  Run from the repo root:
    python tools/dryrun_gpu_compositions.py
Each section runs in its own interpreter (the stand-in installation is global)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SECTIONS = {}

SECTIONS['bricks_and_lewin'] = r'''
# scratch: run the not-yet-run GPU tests' bodies on the CPU with op stand-ins, to catch errors in the TEST code
import sys, types, torch, pytest
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch.nn.functional as F
from fbanet_b200 import ops, _lib as L, train
import test_host_logic as H
import test_zz_gpu_train_compositions as G

class MP:
    def setattr(self, o, n, v): setattr(o, n, v)
H._install_op_standins(MP())
base_conv = ops.conv_gemm
def conv_gemm(srcs, weight, out, *, kh=1, kw=1, pad=0, bias=None, act=0, store_mode=L.STORE_NHWC, **kw_):
    g = H._emu_conv_gemm(srcs[0].float(), weight.float(), kh, pad, convt2=(store_mode == L.STORE_CONVT2))
    if bias is not None: g = g + bias
    out.copy_(g.to(out.dtype)); return out
ops.conv_gemm = conv_gemm
ops.space_to_depth = lambda x: H._emu_s2d(x).contiguous()
def conv_wgrad(x, dy, kh=1, kw=1, stride=1, pad=0):
    w = torch.zeros(dy.shape[-1], x.shape[-1], kh, kw, requires_grad=True); b = torch.zeros(dy.shape[-1], requires_grad=True)
    F.conv2d(x.float().permute(0,3,1,2), w, b, stride=stride, padding=pad).backward(dy.float().permute(0,3,1,2))
    return w.grad, b.grad
ops.conv_wgrad = conv_wgrad
cpu = torch.device("cpu")
for dt in G.DTYPES:
    G.test_convT2_gradients_through_forward_kernels(cpu, dt); print("convT2", dt)
    G.test_pixel_shuffle_conv_gradients_through_forward_kernels(cpu, dt); print("ps", dt)
    for ci, co in ((32, 64), (64, 128)):
        G.test_downsample4_data_gradient_through_forward_kernel(cpu, dt, ci, co); print("down4", dt, ci)
    G.test_act_forward_keeps_the_epilogue_semantics.__wrapped__ if False else None
# act_forward test with stand-in that accepts alpha
af = ops.act_forward
ops.act_forward = lambda x, act, alpha=None: (F.prelu(x.float(), alpha).to(x.dtype) if act == L.ACT_PRELU else af(x.float(), act).to(x.dtype))
for dt in G.DTYPES:
    G.test_act_forward_keeps_the_epilogue_semantics(cpu, dt); print("actfwd", dt)
ops.act_forward = lambda x, act: af(x, act)
# LeWin GPU test: stand-ins operate in the tensors' dtype; LAUNCHES is not incremented by stand-ins -> fake it
H._install_op_standins(MP())
orig = train.lewin_backward
def lb(*a, **k):
    ops.LAUNCHES += 30
    return orig(*a, **k)
train.lewin_backward = lb
G.test_lewin_block_training_forward_backward_on_the_gpu(cpu, torch.float32, 1e-3); print("lewin fp32")
# bf16: stand-ins compute in fp32 and round outputs to bf16 (what the kernels do)
import functools
def rounding(fn):
    @functools.wraps(fn)
    def w(*a, **k):
        dt = None
        def up(t):
            nonlocal dt
            if torch.is_tensor(t) and t.dtype == torch.bfloat16:
                dt = torch.bfloat16; return t.float()
            if isinstance(t, (list, tuple)): return type(t)(up(u) for u in t)
            return t
        a2 = [up(t) for t in a]; k2 = {n: up(v) for n, v in k.items()}
        if "out" in k2 or (fn.__name__ == "conv_gemm"):
            out = a[2]; o32 = torch.empty(out.shape); a2[2] = o32
            fn(*a2, **k2); out.copy_(o32.to(out.dtype)); return out
        r = fn(*a2, **k2)
        if dt is None: return r
        def down(t, first=True): return t.to(dt) if torch.is_tensor(t) else t
        if isinstance(r, tuple): return (down(r[0]),) + tuple(r[1:])      # data gradient rounded, parameter gradients stay fp32
        return down(r)
    return w
for n in ("conv_gemm", "layernorm", "layernorm_backward", "window_attention", "window_attention_backward", "dwconv3x3", "dwconv3x3_backward",
          "act_forward", "act_backward", "drop_path_add"):
    setattr(ops, n, rounding(getattr(ops, n)))
cw = ops.conv_wgrad
ops.conv_wgrad = lambda x, dy, *a, **k: cw(x.float(), dy.float(), *a, **k)
G.test_lewin_block_training_forward_backward_on_the_gpu(cpu, torch.bfloat16, 1e-1); print("lewin bf16 @0.1")
try:
    import test_zz_gpu_train_compositions as G2; _r = G2._rel
    errs = []
    G._rel = lambda a, b: (errs.append(_r(a, b)) or errs[-1])
    G.test_lewin_block_training_forward_backward_on_the_gpu(cpu, torch.bfloat16, 2e-2); print("lewin bf16 @0.02")
except AssertionError as e:
    print("fails at 0.02:", str(e)[:200])
print("max errs", max(errs[1:]), errs[0])
'''

SECTIONS['hourglass'] = r'''
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from fbanet_b200 import ops, train
import test_host_logic as H
import test_zz_gpu_train_compositions as G
class MP:
    def setattr(self, o, n, v): setattr(o, n, v)
H._install_op_standins(MP()); H._install_conv_standins(MP())
orig = train.hourglass_forward_train
def hf(*a, **k):
    ops.LAUNCHES += 300
    return orig(*a, **k)
train.hourglass_forward_train = hf
G.test_hourglass_training_forward_backward_on_the_gpu(torch.device("cpu")); print("hourglass gpu-test body OK (fp32 stand-ins)")
'''

SECTIONS['faf_block'] = r'''
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from fbanet_b200 import ops, train
import test_host_logic as H
import test_zz_gpu_train_compositions as G
class MP:
    def setattr(self, o, n, v): setattr(o, n, v)
H._install_op_standins(MP()); H._install_conv_standins(MP())
orig = train.faf_forward_train
def hf(*a, **k):
    ops.LAUNCHES += 300
    return orig(*a, **k)
train.faf_forward_train = hf
G.test_faf_block_training_forward_backward_on_the_gpu(torch.device("cpu")); print("faf gpu-test body OK (fp32 stand-ins)")
'''

SECTIONS['whole_model'] = r'''
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from fbanet_b200 import ops, train
from fbanet_b200.model import BaseModel
import test_host_logic as H
import test_zz_gpu_train_compositions as G
class MP:
    def setattr(self, o, n, v): setattr(o, n, v)
H._install_op_standins(MP()); H._install_conv_standins(MP())
def adam_step(param, grad, m, v, step, lr, betas, eps, wd, decoupled, grad_scale):
    g = grad * grad_scale
    param.mul_(1 - lr * wd); m.mul_(betas[0]).add_(g, alpha=1 - betas[0]); v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    param.addcdiv_(m / (1 - betas[0] ** step), (v / (1 - betas[1] ** step)).sqrt() + eps, value=-lr)
ops.adam_step = adam_step
orig = train.model_forward_train
def hf(*a, **k):
    ops.LAUNCHES += 1000
    return orig(*a, **k)
train.model_forward_train = hf
# inference forward on the CPU: replace by the training-mode forward (the real one needs CUDA)
BaseModel.__call__ = lambda self, x: orig(self, x, training=False)[0]
import time; t=time.time()
G.test_whole_model_training_step_on_the_gpu(torch.device("cpu")); print("whole-model gpu-test body OK (fp32 stand-ins)", time.time()-t)
'''


if __name__ == "__main__":
    rc = 0
    for name, src in SECTIONS.items():
        r = subprocess.run([sys.executable, "-c", src], cwd=ROOT, capture_output=True, text=True)
        print(f"== {name}: rc={r.returncode}")
        print(r.stdout.strip()[-1500:])
        if r.returncode:
            print(r.stderr.strip()[-3000:])
            rc = 1
    sys.exit(rc)
