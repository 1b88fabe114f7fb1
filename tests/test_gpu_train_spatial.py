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


# ---- resampling layers: gradients as compositions of the FORWARD kernels with re-packed weights (train.dgrad_weight_* / *_wgrad).
# The packings are checked on the CPU against an emulation of the kernels' contracts (tests/test_host_logic.py::
# test_resampling_layer_gradients_through_forward_kernels); here the real kernels run them.
@pytest.mark.parametrize("dtype", DTYPES)
def test_convT2_gradients_through_forward_kernels(cuda, dtype):
    """ConvTranspose2d(4E, E, 2, 2) of layers/upsample.py:19-30 (E = 64)."""
    from fbanet_b200 import train
    g = torch.Generator().manual_seed(21)
    N, H, W, Ci, Co = 2, 10, 12, 256, 64
    w = (torch.randn(Ci, Co, 2, 2, generator=g) / Ci ** 0.5).to(cuda)
    x = torch.randn(N, H, W, Ci, generator=g).to(dtype).to(cuda)
    dy = torch.randn(N, 2 * H, 2 * W, Co, generator=g).to(dtype).to(cuda)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    wr = w.to(dtype).float().requires_grad_(True)
    br = torch.zeros(Co, device=cuda, requires_grad=True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        F.conv_transpose2d(xr, wr, br, stride=2).backward(dy.float().permute(0, 3, 1, 2))
    dx = train.convT2_dgrad(dy, w)
    assert _rel(dx.float(), xr.grad.permute(0, 2, 3, 1)) < _tol(dtype), _rel(dx.float(), xr.grad.permute(0, 2, 3, 1))
    dw, db = train.convT2_wgrad(x, dy)
    assert dw.shape == w.shape and _rel(dw, wr.grad) < 2e-4 and _rel(db, br.grad) < 2e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_pixel_shuffle_conv_gradients_through_forward_kernels(cuda, dtype):
    """conv3x3 E -> 4E + PixelShuffle(2) of blocks/upsampler.py:22-32 (E = 64)."""
    from fbanet_b200 import train
    g = torch.Generator().manual_seed(23)
    N, H, W, Ci, C = 2, 10, 12, 64, 64
    w = (torch.randn(4 * C, Ci, 3, 3, generator=g) / (9 * Ci) ** 0.5).to(cuda)
    x = torch.randn(N, H, W, Ci, generator=g).to(dtype).to(cuda)
    dy = torch.randn(N, 2 * H, 2 * W, C, generator=g).to(dtype).to(cuda)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    wr = w.to(dtype).float().requires_grad_(True)
    br = torch.zeros(4 * C, device=cuda, requires_grad=True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        F.pixel_shuffle(F.conv2d(xr, wr, br, padding=1), 2).backward(dy.float().permute(0, 3, 1, 2))
    dx = train.pixel_shuffle_conv_dgrad(dy, w)
    assert _rel(dx.float(), xr.grad.permute(0, 2, 3, 1)) < _tol(dtype), _rel(dx.float(), xr.grad.permute(0, 2, 3, 1))
    dw, db = train.pixel_shuffle_conv_wgrad(x, dy)
    assert dw.shape == w.shape and _rel(dw, wr.grad) < 2e-4 and _rel(db, br.grad) < 2e-4


@pytest.mark.parametrize("dtype,Ci,Co", [(torch.float32, 32, 64), (torch.float32, 64, 128), (torch.bfloat16, 32, 64), (torch.bfloat16, 64, 128)])
def test_downsample4_data_gradient_through_forward_kernel(cuda, dtype, Ci, Co):
    """Conv2d(E, 2E, 4, 2, 1) of layers/downsample.py:19-30 (E = 32, 64): four sub-pixel phases as rows of one 3x3 GEMM over dy."""
    from fbanet_b200 import train
    g = torch.Generator().manual_seed(22)
    N, Ho, Wo = 2, 10, 12
    w = (torch.randn(Co, Ci, 4, 4, generator=g) / (16 * Ci) ** 0.5).to(cuda)
    dy = torch.randn(N, Ho, Wo, Co, generator=g).to(dtype).to(cuda)
    xr = torch.zeros(N, Ci, 2 * Ho, 2 * Wo, device=cuda, requires_grad=True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        F.conv2d(xr, w.to(dtype).float(), None, stride=2, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    dx = train.down4_dgrad(dy, w)
    assert dx.shape == (N, 2 * Ho, 2 * Wo, Ci)
    assert _rel(dx.float(), xr.grad.permute(0, 2, 3, 1)) < _tol(dtype), _rel(dx.float(), xr.grad.permute(0, 2, 3, 1))


@pytest.mark.parametrize("dtype", DTYPES)
def test_act_forward_keeps_the_epilogue_semantics(cuda, dtype):
    """Stand-alone training-mode activation == the functions the GEMM epilogues apply (common.cuh apply_act)."""
    from fbanet_b200 import _lib as L, ops
    x = (torch.randn(3, 11, 13, 24, generator=torch.Generator().manual_seed(4)) * 2).to(dtype).to(cuda)
    alpha = torch.tensor([0.25], device=cuda)
    for act, fn in ((L.ACT_NONE, lambda v: v), (L.ACT_RELU, F.relu), (L.ACT_GELU_TANH, lambda v: F.gelu(v, approximate="tanh")),
                    (L.ACT_GELU_ERF, F.gelu), (L.ACT_PRELU, lambda v: F.prelu(v, alpha))):
        y = ops.act_forward(x, act, alpha=alpha if act == L.ACT_PRELU else None)
        assert _rel(y.float(), fn(x.float())) < (1e-6 if dtype == torch.float32 else 4e-3), act


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 1e-1)])
def test_lewin_block_training_forward_backward_on_the_gpu(cuda, dtype, tol):
    """train.lewin_forward_train / lewin_backward (the composition tests/test_host_logic.py checks with op stand-ins) on the real
    kernels, against autograd through the oracle's LeWinLayer in float64 on the CPU (layers/fba_net.py:139-250, Appendix A-4);
    fp32 is the parity path, bf16 rounds every saved activation (loose bound, relative to each gradient's largest entry)."""
    from fbanet_b200 import train, ops
    from fbanet_b200.model import _Layer
    from oracle.fbanet_oracle import LeWinLayer
    torch.manual_seed(5)
    dim, res, heads, win, shift, B = 64, (20, 20), 2, 10, 5, 2
    ly = _Layer(dim, res, heads, win, shift, 4.0)
    with torch.no_grad():
        for n, p in ly.named_parameters():
            p.copy_(torch.randn_like(p) * (0.5 if p.dim() == 1 or "table" in n else 1.5 / p[0].numel() ** 0.5))
            if n.endswith(("norm1.weight", "norm2.weight")):
                p.add_(1.0)
    ref = LeWinLayer(dim, res, heads, win, shift, 4.0, "tanh").double()
    ref.load_state_dict(ly.state_dict())
    x = torch.randn(B, *res, dim).to(dtype)
    dy = torch.randn(B, *res, dim).to(dtype)
    xr = x.double().requires_grad_(True)
    yr = ref(xr.view(B, -1, dim)).view(B, *res, dim)
    yr.backward(dy.double())
    ly = ly.to(cuda)
    before = ops.LAUNCHES
    y, saved = train.lewin_forward_train(ly, x.to(cuda))
    dx = train.lewin_backward(ly, saved, dy.to(cuda))
    assert ops.LAUNCHES - before >= 30                                   # the CUDA ops ran (no torch fallback inside the composition)
    assert _rel(y.float().cpu(), yr.detach()) < (2e-4 if dtype == torch.float32 else 4e-2), _rel(y.float().cpu(), yr.detach())
    assert _rel(dx.float().cpu(), xr.grad) < tol, _rel(dx.float().cpu(), xr.grad)
    got = dict(ly.named_parameters())
    for n, pr in ref.named_parameters():
        assert got[n].grad is not None and _rel(got[n].grad.cpu(), pr.grad) < tol, (n, _rel(got[n].grad.cpu(), pr.grad))


def test_hourglass_training_forward_backward_on_the_gpu(cuda):
    """train.hourglass_forward_train + Tape.backward on the real kernels (fp32 parity path): ten LeWin layers, 4x4 s2 downsamples,
    transposed-conv upsamples, skip concats -- against autograd through the oracle's first hourglass in float64 on the CPU
    (models/fba_net.py:271-287).  The composition itself is checked with op stand-ins in tests/test_host_logic.py."""
    from fbanet_b200 import train, ops
    from fbanet_b200.model import BaseModel
    from oracle.fbanet_oracle import OracleBaseModel
    cfg = dict(num_frames=2, img_size=40, embed_dim=32, window_length=10)
    m = BaseModel(token_mlp="leff", dtype="fp32", seed=1, **cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "relative_position_bias_table" in n or (p.dim() == 1 and "norm" not in n):
                p.copy_(torch.randn_like(p) * 0.2)
    o = OracleBaseModel(**cfg).double()
    assert not o.load_state_dict(m.state_dict(), strict=False).missing_keys
    B, S, E = 2, 40, 32
    y = torch.randn(B, S, S, E, generator=torch.Generator().manual_seed(2))
    dout = torch.randn(B, S, S, 2 * E, generator=torch.Generator().manual_seed(3))
    yr = y.double().requires_grad_(True)
    ref, _ = o._hourglass("HG1", yr.view(B, S * S, E))
    ref.backward(dout.double().view(B, S * S, 2 * E))
    m = m.to(cuda)
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    yd = y.to(cuda)
    before = ops.LAUNCHES
    out, tape = train.hourglass_forward_train(m, "HG1", yd, training=False)
    grads = tape.backward(out, dout.to(cuda))
    assert ops.LAUNCHES - before >= 10 * 30
    assert _rel(out.cpu().view(B, S * S, 2 * E), ref.detach()) < 5e-4
    assert set(grads) == {id(yd)} and _rel(grads[id(yd)].cpu(), yr.grad.view(B, S, S, E)) < 2e-3
    got, n_checked = dict(m.named_parameters()), 0
    for n, pr in o.named_parameters():
        if pr.grad is None:
            continue
        assert got[n].grad is not None and _rel(got[n].grad.cpu(), pr.grad) < 2e-3, (n, _rel(got[n].grad.cpu(), pr.grad))
        n_checked += 1
    assert n_checked == 10 * 17 + 4 * 2


def test_faf_block_training_forward_backward_on_the_gpu(cuda):
    """train.faf_forward_train + Tape.backward on the real kernels (fp32 parity path): gate, 1x1 fusion + PReLU, ten ResBlocks with
    4x4 s2 / transposed-conv resampling, fusion_tail, skip -- against autograd through the oracle's FAFBlock as written, float64 on
    the CPU (blocks/federated_affinity_fusion.py:166-182).  Composition checked with op stand-ins in tests/test_host_logic.py."""
    from fbanet_b200 import train, ops
    from fbanet_b200.model import _FAF
    from oracle.fbanet_oracle import FAFBlock
    torch.manual_seed(12)
    E, Fr, B, S = 32, 3, 2, 16
    fu = _FAF(E, Fr)
    with torch.no_grad():
        for p in fu.parameters():
            p.copy_(torch.randn_like(p) * (0.3 if p.dim() == 1 else 0.7 / p[0].numel() ** 0.5))
    ref = FAFBlock(E, Fr).double()
    ref.load_state_dict(fu.state_dict())
    feat = torch.randn(B, Fr, S, S, E)
    dout = torch.randn(B, S, S, E)
    fr = feat.double().requires_grad_(True)
    yr = ref(fr.permute(0, 1, 4, 2, 3))
    yr.backward(dout.double().permute(0, 3, 1, 2))
    fu = fu.to(cuda)
    fd = feat.to(cuda)
    before = ops.LAUNCHES
    out, tape = train.faf_forward_train(fu, fd)
    grads = tape.backward(out, dout.to(cuda))
    assert ops.LAUNCHES - before >= 100
    assert _rel(out.cpu(), yr.detach().permute(0, 2, 3, 1)) < 5e-4
    assert set(grads) == {id(fd)} and _rel(grads[id(fd)].cpu(), fr.grad) < 2e-3, _rel(grads[id(fd)].cpu(), fr.grad)
    got = dict(fu.named_parameters())
    for n, pr in ref.named_parameters():
        if n.startswith("temporal_attn0") or n == "temporal_attn1.bias":
            assert got[n].grad is None, n
            continue
        assert got[n].grad is not None and _rel(got[n].grad.cpu(), pr.grad) < 2e-3, (n, _rel(got[n].grad.cpu(), pr.grad))


def test_whole_model_training_step_on_the_gpu(cuda):
    """BASELINE config 5 on the real kernels (fp32 parity path, small model): (1) the training-mode forward of the whole model
    reproduces the inference forward; (2) loss + Tape.backward give the gradients autograd gives through the oracle model and the
    oracle's CharbonnierLoss + 3 GWLoss (float64, CPU) for EVERY parameter; (3) two train_step calls (flat buffers, fused AdamW)
    on the same sample lower the loss.  The composition is checked with op stand-ins in tests/test_host_logic.py."""
    from fbanet_b200 import train, ops
    from fbanet_b200.model import BaseModel
    from oracle.fbanet_oracle import OracleBaseModel, training_loss
    cfg = dict(num_frames=3, img_size=40, embed_dim=32, window_length=10)
    m = BaseModel(token_mlp="leff", dtype="fp32", seed=4, **cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "relative_position_bias_table" in n or (p.dim() == 1 and "norm" not in n):
                p.copy_(torch.randn_like(p) * 0.1)
    o = OracleBaseModel(**cfg).double()
    o.load_state_dict(m.state_dict())
    B = 2
    burst = torch.rand(B, 3, 3, 40, 40, generator=torch.Generator().manual_seed(1))
    target = torch.rand(B, 3, 160, 160, generator=torch.Generator().manual_seed(2))
    ref = o(burst.double())
    loss_ref = training_loss(ref, target.double())
    loss_ref.backward()
    m = m.to(cuda)
    bd, td = burst.to(cuda), target.to(cuda)
    infer = m(bd)                                                          # the inference path (graph of fused kernels)
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    before = ops.LAUNCHES
    restored, tape = train.model_forward_train(m, bd, training=False)
    assert _rel(restored, infer) < 1e-4 and _rel(restored.cpu(), ref.detach()) < 1e-3
    loss, d_restored = ops.training_loss(restored, td)
    assert abs(loss[0].item() - loss_ref.item()) < 1e-4 * loss_ref.item()
    assert tape.backward(restored, d_restored) == {}
    assert ops.LAUNCHES - before >= 20 * 30 + 100                          # the C-ABI ops ran: no torch arithmetic on the path
    got, worst = dict(m.named_parameters()), 0.0
    for n, pr in o.named_parameters():
        if n.startswith("fusion.temporal_attn0") or n == "fusion.temporal_attn1.bias":
            assert got[n].grad is None, n
            continue
        assert got[n].grad is not None, n
        e = _rel(got[n].grad.cpu(), pr.grad)
        worst = max(worst, e)
        assert e < 5e-3, (n, e)
    # the optimizer loop
    for p in m.parameters():
        p.grad = None
    m.drop_path_rate = 0.0
    flat = train.FlatParams(m.parameters())
    l1 = train.train_step(m, flat, bd, td, lr=2e-4)
    l2 = train.train_step(m, flat, bd, td, lr=2e-4)
    l3 = train.train_step(m, flat, bd, td, lr=2e-4)
    assert flat.step == 3 and torch.isfinite(l3).all() and l3[0].item() < l1[0].item(), (l1, l2, l3)
