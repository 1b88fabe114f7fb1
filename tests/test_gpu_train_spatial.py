"""Second set of backward bricks of the training step (depthwise 3x3, window attention, Federated-Affinity gate,
DropPath residual) against torch autograd of the same op on the same inputs (train.py.bak:163-169 trains through torch autograd;
the forward expressions below are the ones tests/test_gpu_ops.py pins the forward kernels to).  Tolerances: 2e-4 of the largest
gradient entry for fp32 inputs; bf16 inputs are compared against autograd on the SAME bf16-rounded values upcast to fp32, so only
the output rounding (2^-8) and the accumulation order differ."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


def _tol(dtype):
    return 2e-4 if dtype == torch.float32 else 1e-2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,H,W,C", [(2, 14, 10, 256), (1, 7, 9, 70), (3, 5, 5, 64)])
def test_dwconv3x3_backward_matches_autograd(cuda, dtype, N, H, W, C):
    """layers/locally_enhanced_feed_forward.py:39-44: depthwise conv, groups = hidden, k3, p1."""
    from fbanet_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, H, W, C, generator=g).to(dtype).to(cuda)
    dy = torch.randn(N, H, W, C, generator=g).to(dtype).to(cuda)
    w = (torch.randn(C, 1, 3, 3, generator=g) * 0.3).to(cuda).requires_grad_(True)
    b = torch.zeros(C, device=cuda, requires_grad=True)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        F.conv2d(xr, w, b, padding=1, groups=C).backward(dy.float().permute(0, 3, 1, 2))
    w9c = w.detach().reshape(C, 9).t().contiguous()
    dx, dw, db = ops.dwconv3x3_backward(x, dy, w9c)
    assert _rel(dx.float().permute(0, 3, 1, 2), xr.grad) < _tol(dtype)
    assert dw.shape == w.shape and _rel(dw, w.grad) < 2e-4 and _rel(db, b.grad) < 2e-4
    dx2, dw2, db2 = ops.dwconv3x3_backward(x, dy, w9c)
    assert torch.equal(dx, dx2) and torch.equal(dw, dw2) and torch.equal(db, db2)          # fixed-order reductions
    ops.dwconv3x3_backward(x, dy, w9c, need_dx=False, dw=dw2, db=db2, accumulate=True)
    assert _rel(dw2, 2 * w.grad) < 2e-4 and _rel(db2, 2 * b.grad) < 2e-4


def _attention_reference(qkv, table, B, H, W, C, heads, win, shift, scale):
    """The forward expression of tests/test_gpu_ops.py::test_window_attention (oracle helpers inlined so that it runs on the GPU)."""
    N, dh = win * win, C // heads
    y = qkv.view(B, H, W, 3 * C)
    if shift:
        y = torch.roll(y, (-shift, -shift), (1, 2))
    yw = y.view(B, H // win, win, W // win, win, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, N, 3 * C)
    q, k, v = (yw[..., i * C:(i + 1) * C].reshape(-1, N, heads, dh).permute(0, 2, 1, 3) for i in range(3))
    attn = (q * scale) @ k.transpose(-2, -1)
    co = torch.stack(torch.meshgrid(torch.arange(win), torch.arange(win), indexing="ij")).flatten(1)
    rel = co[:, :, None] - co[:, None, :] + (win - 1)
    idx = (rel[0] * (2 * win - 1) + rel[1]).to(qkv.device)
    attn = attn + table[idx.view(-1)].view(N, N, heads).permute(2, 0, 1)[None]
    if shift:
        img = torch.zeros(H, W)
        cnt = 0
        for hs in (slice(0, -win), slice(-win, -shift), slice(-shift, None)):
            for ws in (slice(0, -win), slice(-win, -shift), slice(-shift, None)):
                img[hs, ws] = cnt
                cnt += 1
        mw = img.view(H // win, win, W // win, win).permute(0, 2, 1, 3).reshape(-1, N)
        mask = (mw[:, None, :] - mw[:, :, None]).ne(0).float().mul(-100.0).to(qkv.device)
        attn = (attn.view(B, mask.shape[0], heads, N, N) + mask[None, :, None]).view(-1, heads, N, N)
    o = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(B, H // win, W // win, win, win, C)
    o = o.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o.reshape(-1, C)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,heads,HW,win,shift", [(64, 1, 20, 10, 0), (64, 1, 20, 10, 5), (128, 2, 20, 10, 5), (256, 16, 20, 10, 5),
                                                   (128, 8, 8, 4, 2), (64, 4, 5, 5, 0)])
def test_window_attention_backward_matches_autograd(cuda, dtype, C, heads, HW, win, shift):
    """layers/window_attention.py:159-248 + shift / partition / mask of layers/fba_net.py:149-238."""
    from fbanet_b200 import ops
    B, H, W = 2, HW, HW
    g = torch.Generator().manual_seed(7)
    qkv = torch.randn(B * H * W, 3 * C, generator=g).to(dtype).to(cuda)
    dout = torch.randn(B * H * W, C, generator=g).to(dtype).to(cuda)
    table = (torch.randn((2 * win - 1) ** 2, heads, generator=g) * 0.5).to(cuda)
    scale = (C // heads) ** -0.5
    # the forward kernel agrees with the expression the gradient is taken of
    out = ops.window_attention(qkv, table, B, H, W, heads, win, shift, scale)
    qr, tr = qkv.float().requires_grad_(True), table.clone().requires_grad_(True)
    ref = _attention_reference(qr, tr, B, H, W, C, heads, win, shift, scale)
    assert _rel(out.float(), ref.detach()) < (1e-4 if dtype == torch.float32 else 3e-2)
    ref.backward(dout.float())
    dqkv, dbias = ops.window_attention_backward(qkv, dout, table, B, H, W, heads, win, shift, scale)
    for i, name in enumerate("qkv"):
        a, b = dqkv[:, i * C:(i + 1) * C].float(), qr.grad[:, i * C:(i + 1) * C]
        assert _rel(a, b) < _tol(dtype), (name, _rel(a, b))
    assert dbias.shape == table.shape and _rel(dbias, tr.grad) < 2e-4, _rel(dbias, tr.grad)
    dqkv2, dbias2 = ops.window_attention_backward(qkv, dout, table, B, H, W, heads, win, shift, scale)
    assert torch.equal(dqkv, dqkv2) and torch.equal(dbias, dbias2)
    ops.window_attention_backward(qkv, dout, table, B, H, W, heads, win, shift, scale, dbias=dbias2, accumulate=True)
    assert _rel(dbias2, 2 * tr.grad) < 2e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,Fr,E,S", [(2, 6, 64, 14), (1, 3, 32, 9)])
def test_faf_gate_backward_matches_autograd(cuda, dtype, B, Fr, E, S):
    """blocks/federated_affinity_fusion.py:79-105 AS WRITTEN (two embedding convs, sum over channels of emb - emb_ref, gate from
    |aff_f - aff_0|): gradients into the features and into BOTH convolutions' weights and biases (temporal_attn0 and the biases get 0)."""
    from fbanet_b200 import ops
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(B, Fr, E, S, S, generator=g).to(dtype).to(cuda)
    dgated = torch.randn(B, S, S, Fr * E, generator=g).to(dtype).to(cuda)
    w0 = (torch.randn(E, E, 3, 3, generator=g) * 0.01).to(cuda).requires_grad_(True)
    w1 = (torch.randn(E, E, 3, 3, generator=g) * 0.01).to(cuda).requires_grad_(True)
    b0 = (torch.randn(E, generator=g) * 0.1).to(cuda).requires_grad_(True)
    b1 = (torch.randn(E, generator=g) * 0.1).to(cuda).requires_grad_(True)
    fr = feat.float().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        ref_e = F.conv2d(fr[:, 0], w0, b0, padding=1)
        emb = F.conv2d(fr.reshape(B * Fr, E, S, S), w1, b1, padding=1).view(B, Fr, E, S, S)
    aff = (emb - ref_e[:, None]).sum(2)
    gate_ref = torch.sigmoid((aff[:, 1:] - aff[:, :1]).abs())
    gated = torch.cat([fr[:, :1], fr[:, 1:] * gate_ref[:, :, None]], 1)                 # [B,F,E,S,S]
    gated.permute(0, 3, 4, 1, 2).reshape(B, S, S, Fr * E).backward(dgated.float())
    # kernel inputs: channels-last features, the forward kernel's own gate, fp32 scores s_f = wsum (*) feat_f
    fd = feat.permute(0, 1, 3, 4, 2).contiguous()
    wsum = w1.detach().sum(0).permute(1, 2, 0).reshape(9, E).contiguous()
    gate = ops.faf_gate(fd, wsum)
    assert (gate - gate_ref.detach()).abs().max().item() < 1e-4
    with torch.backends.cudnn.flags(allow_tf32=False):
        score = F.conv2d(feat.float().reshape(B * Fr, E, S, S), w1.detach().sum(0, keepdim=True), None, padding=1).reshape(B * Fr, S, S)
    dfeat, dwsum = ops.faf_gate_backward(fd, dgated, gate, score, wsum)
    assert _rel(dfeat.float().permute(0, 1, 4, 2, 3), fr.grad) < _tol(dtype), _rel(dfeat.float().permute(0, 1, 4, 2, 3), fr.grad)
    dw1 = ops.faf_weight_grads(dwsum, E)
    assert dw1.shape == w1.shape and _rel(dw1, w1.grad) < 5e-4, _rel(dw1, w1.grad)
    # emb_ref and both biases cancel out of aff_f - aff_0: their gradients vanish (up to the rounding of the cancellation)
    lim = 1e-4 * w1.grad.abs().max().item()
    assert w0.grad.abs().max().item() < lim and b0.grad.abs().max().item() < lim * S * S and b1.grad.abs().max().item() < lim * S * S
    dfeat2, dwsum2 = ops.faf_gate_backward(fd, dgated, gate, score, wsum)
    assert torch.equal(dfeat, dfeat2) and torch.equal(dwsum, dwsum2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_drop_path_add_forward_and_backward(cuda, dtype):
    """layers/drop_path.py:52-63 ("global" mode, one draw per burst) inside the residuals of layers/fba_net.py:245,248."""
    from fbanet_b200 import ops, train
    B, T, C = 5, 37, 64
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, C, generator=g).to(dtype).to(cuda)
    skip = torch.randn(B, T, C, generator=g).to(dtype).to(cuda)
    scale = train.drop_path_scales(B, 0.4, generator=torch.Generator().manual_seed(2), device=cuda)
    assert all(v == 0.0 or abs(v - 1.0 / 0.6) < 1e-6 for v in scale.tolist()) and 0.0 in scale.tolist() and scale.max().item() > 1.0
    x[scale == 0] = float("inf")                                                  # a dropped branch contributes exactly nothing
    out = ops.drop_path_add(x, scale, skip)
    want = skip.float() + torch.where(scale[:, None, None] == 0, torch.zeros_like(x.float()), x.float() * scale[:, None, None])
    assert torch.equal(out, want.to(dtype))
    dx = ops.drop_path_add(skip, scale)                                            # backward of the branch: dy * scale
    assert torch.equal(dx, (skip.float() * scale[:, None, None]).to(dtype))
    ones = train.drop_path_scales(B, 0.0, device=cuda)
    assert torch.equal(ops.drop_path_add(skip, ones), skip)
