"""Backward bricks of the training step against torch autograd on the same inputs -- for these floating-point ops
autograd of the fp32 torch op IS the reference's gradient (train.py.bak:163-169 trains through torch autograd).  Tolerances: fp32
inputs 2e-4 relative to the largest gradient entry (sums over up to 10^5 pixels in fp32); bf16 inputs are compared against autograd
on the SAME bf16-rounded values upcast to fp32, so only the accumulation differs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [
    dict(N=2, H=20, W=24, Cin=64, Cout=64, k=3, s=1, p=1),      # body / FAF 3x3
    dict(N=1, H=40, W=40, Cin=128, Cout=512, k=1, s=1, p=0),    # fc1 (a linear layer over tokens)
    dict(N=2, H=16, W=12, Cin=64, Cout=128, k=4, s=2, p=1),     # downsample 4x4 stride 2
    dict(N=3, H=9, W=7, Cin=3, Cout=20, k=3, s=1, p=1),         # ragged channels (head conv), odd sizes
    dict(N=1, H=10, W=10, Cin=70, Cout=130, k=3, s=1, p=1),     # channel tiles with remainders
])
def test_conv_wgrad_matches_autograd(cuda, dtype, case):
    from fbanet_b200 import ops
    c = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(c["N"], c["H"], c["W"], c["Cin"], generator=g).to(dtype).to(cuda)
    Ho, Wo = (c["H"] + 2 * c["p"] - c["k"]) // c["s"] + 1, (c["W"] + 2 * c["p"] - c["k"]) // c["s"] + 1
    dy = torch.randn(c["N"], Ho, Wo, c["Cout"], generator=g).to(dtype).to(cuda)
    dw, db = ops.conv_wgrad(x, dy, c["k"], c["k"], c["s"], c["p"])
    w = torch.zeros(c["Cout"], c["Cin"], c["k"], c["k"], device=cuda, requires_grad=True)
    b = torch.zeros(c["Cout"], device=cuda, requires_grad=True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        y = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, stride=c["s"], padding=c["p"])
        y.backward(dy.float().permute(0, 3, 1, 2))
    assert dw.shape == w.shape and _rel(dw, w.grad) < 2e-4, _rel(dw, w.grad)
    assert _rel(db, b.grad) < 2e-4, _rel(db, b.grad)
    # reproducible, and `accumulate` adds
    dw2, db2 = ops.conv_wgrad(x, dy, c["k"], c["k"], c["s"], c["p"])
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    ops.conv_wgrad(x, dy, c["k"], c["k"], c["s"], c["p"], dw=dw2, db=db2, accumulate=True)
    assert _rel(dw2, 2 * w.grad) < 2e-4 and _rel(db2, 2 * b.grad) < 2e-4


def test_conv_wgrad_reads_channel_slices(cuda):
    """x and dy as channel slices of wider buffers (the concat-free skip-connection layout of DESIGN.md 3) and an explicit split count."""
    from fbanet_b200 import ops
    g = torch.Generator().manual_seed(5)
    xb = torch.randn(2, 12, 12, 96, generator=g).to(cuda)
    dyb = torch.randn(2, 12, 12, 80, generator=g).to(cuda)
    x, dy = xb[..., 32:96], dyb[..., 8:40]
    dw, db = ops.conv_wgrad(x, dy, 3, 3, 1, 1, splits=5)
    w = torch.zeros(32, 64, 3, 3, device=cuda, requires_grad=True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        F.conv2d(x.permute(0, 3, 1, 2), w, None, padding=1).backward(dy.permute(0, 3, 1, 2))
    assert _rel(dw, w.grad) < 2e-4
    assert _rel(db, dy.sum((0, 1, 2))) < 2e-4


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("k,pad", [(3, 1), (1, 0)])
def test_conv_dgrad_is_the_forward_gemm_with_flipped_weights(cuda, dtype, tol, k, pad):
    """Data gradient of a stride-1 conv / linear layer through the FORWARD kernel (`train.dgrad_weight`); bf16 -> tcgen05 path."""
    from fbanet_b200 import ops
    from fbanet_b200.train import dgrad_weight
    g = torch.Generator().manual_seed(3)
    N, H, W, Cin, Cout = 2, 20, 20, 64, 128
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(cuda)
    dy = torch.randn(N, H, W, Cout, generator=g).to(dtype).to(cuda)
    dx = torch.empty(N, H, W, Cin, device=cuda, dtype=dtype)
    ops.conv_gemm([dy], dgrad_weight(w, dtype), dx, kh=k, kw=k, pad=k - 1 - pad)
    x = torch.zeros(N, Cin, H, W, device=cuda, requires_grad=True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        F.conv2d(x, w.to(dtype).float(), None, padding=pad).backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(dx.float(), x.grad.permute(0, 2, 3, 1)) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("C", [64, 128, 256, 48])
def test_layernorm_backward_matches_autograd(cuda, dtype, tol, C):
    from fbanet_b200 import ops
    g = torch.Generator().manual_seed(C)
    rows = 2500
    x = (torch.randn(rows, C, generator=g) * 1.5 + 0.3).to(dtype).to(cuda)
    dy = torch.randn(rows, C, generator=g).to(dtype).to(cuda)
    gamma = (torch.rand(C, generator=g) + 0.5).to(cuda)
    beta = torch.randn(C, generator=g).to(cuda)
    dx, dgamma, dbeta = ops.layernorm_backward(x, dy, gamma)
    xr = x.detach().float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (C,), gr, br, 1e-5).backward(dy.float())
    assert _rel(dx.float(), xr.grad) < tol, _rel(dx.float(), xr.grad)   # bf16: dx itself is rounded to bf16
    assert _rel(dgamma, gr.grad) < 2e-4 and _rel(dbeta, br.grad) < 2e-4
    dx2, dgamma2, dbeta2 = ops.layernorm_backward(x, dy, gamma)
    assert torch.equal(dx, dx2) and torch.equal(dgamma, dgamma2) and torch.equal(dbeta, dbeta2)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-5), (torch.bfloat16, 1e-2)])
def test_activation_backward_matches_autograd(cuda, dtype, tol):
    from fbanet_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(3, 50, 70, 16, generator=g) * 2).to(dtype).to(cuda)
    dy = torch.randn(3, 50, 70, 16, generator=g).to(dtype).to(cuda)
    alpha = torch.tensor([0.25], device=cuda)
    for act, fn in ((L.ACT_RELU, F.relu), (L.ACT_GELU_TANH, lambda v: F.gelu(v, approximate="tanh")), (L.ACT_GELU_ERF, F.gelu)):
        xr = x.detach().float().clone().requires_grad_(True)   # a fresh leaf per activation (x.float() aliases x in fp32)
        fn(xr).backward(dy.float())
        assert _rel(ops.act_backward(x, dy, act).float(), xr.grad) < tol, act
    xr, ar = x.detach().float().clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    F.prelu(xr, ar).backward(dy.float())
    dx, dalpha = ops.act_backward(x, dy, L.ACT_PRELU, alpha=alpha)
    assert _rel(dx.float(), xr.grad) < tol and _rel(dalpha, ar.grad) < 2e-4
    _, dalpha2 = ops.act_backward(x, dy, L.ACT_PRELU, alpha=alpha, dalpha=dalpha.clone(), accumulate=True)
    assert _rel(dalpha2, 2 * ar.grad) < 2e-4
