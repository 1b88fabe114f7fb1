#!/usr/bin/env python
"""Time (or feed ncu with) the LayerNorm-in-GEMM kernel on a qkv shape: `python tools/prof_ln_gemm.py [C cout S reps]`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbanet_b200 import ops, _lib as L

C, cout, S, reps = (int(a) for a in (sys.argv[1:5] + ["128", "384", "160", "10"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
BF = torch.bfloat16
x = (torch.rand(64, S, S, C, device=dev) - 0.5).to(BF)
w = ((torch.rand(cout, C, device=dev) - 0.5) * 0.1).to(BF)
b = torch.zeros(cout, device=dev)
g, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
out = torch.empty(64, S, S, cout, device=dev, dtype=BF)
for name, fn in (("ln in gemm", lambda: ops.conv_gemm([x], w, out, bias=b, impl=L.IMPL_TCGEN05, ln=(g, be))),
                 ("plain gemm", lambda: ops.conv_gemm([x], w, out, bias=b, impl=L.IMPL_TCGEN05)),
                 ("layernorm ", lambda: ops.layernorm(x.view(-1, C), g, be))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name} {C}->{cout} @{S}^2 x64: {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)
