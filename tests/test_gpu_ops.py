"""Per-kernel parity: every C-ABI op against the CPU oracle / torch CPU functional ops on seeded inputs.
fp32 path: tight tolerances (fp32 re-association only).  bf16 path: inputs/weights rounded to bf16 on both
sides, tolerance = bf16 output rounding."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def _tol(dt):
    return (2e-5, 2e-5) if dt == torch.float32 else (2e-2, 2e-2)


def _close(got, ref, dt, scale=1.0):
    atol, rtol = _tol(dt)
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs().max().item()
    assert torch.allclose(got, ref, atol=atol * scale, rtol=rtol), f"max abs err {err}"


def _nhwc(x, dt, dev):  # [N,C,H,W] cpu -> [N,H,W,C] device
    return x.permute(0, 2, 3, 1).contiguous().to(device=dev, dtype=dt)


def _pack(w, dt, dev):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(device=dev, dtype=dt)


def _r(dt, *shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = (torch.rand(shape, generator=g) * 2 - 1) * scale
    return t.to(dt).float() if dt == torch.bfloat16 else t


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("cin,cout,hw,act", [(3, 64, 20, 0), (64, 64, 24, 1), (128, 64, 16, 2), (64, 256, 10, 3), (64, 3, 12, 0)])
def test_conv3x3(cuda, dt, cin, cout, hw, act):
    from fbanet_b200 import ops, _lib as L
    x = _r(dt, 2, cin, hw, hw, seed=1)
    w = _r(dt, cout, cin, 3, 3, seed=2, scale=1 / math.sqrt(cin * 9))
    b = _r(torch.float32, cout, seed=3, scale=0.1)
    res = _r(dt, 2, cout, hw, hw, seed=4)
    alpha = torch.tensor([0.25])
    ref = F.conv2d(x, w, b, padding=1)
    ref = {0: lambda v: v, 1: F.relu, 2: lambda v: F.prelu(v, alpha), 3: lambda v: F.gelu(v, approximate="tanh")}[act](ref) + res
    cp = cin if cin % 8 == 0 else 8
    xd = torch.zeros(2, hw, hw, cp, device=cuda, dtype=dt)
    xd[..., :cin] = _nhwc(x, dt, cuda)
    wp = torch.zeros(cout, 3, 3, cp)
    wp[..., :cin] = w.permute(0, 2, 3, 1)
    out = torch.empty(2, hw, hw, cout, device=cuda, dtype=dt)
    ops.conv_gemm([xd], wp.reshape(cout, -1).to(cuda, dt), out, kh=3, kw=3, pad=1, bias=b.to(cuda), act=act,
                  alpha=alpha.to(cuda), residual=_nhwc(res, dt, cuda))
    _close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_multisource_strided_views(cuda, dt):
    """concat-free inputs: three sources (one a channel slice of a wider buffer), output into a slice."""
    from fbanet_b200 import ops
    a, b_, c = _r(dt, 2, 64, 12, 12, seed=1), _r(dt, 2, 32, 12, 12, seed=2), _r(dt, 2, 32, 12, 12, seed=3)
    w = _r(dt, 64, 128, 3, 3, seed=4, scale=0.03)
    bias = _r(torch.float32, 64, seed=5)
    ref = F.conv2d(torch.cat([a, b_, c], 1), w, bias, padding=1)
    wide = torch.zeros(2, 12, 12, 96, device=cuda, dtype=dt)
    wide[..., 64:] = _nhwc(c, dt, cuda)
    outbuf = torch.zeros(2, 12, 12, 128, device=cuda, dtype=dt)
    ops.conv_gemm([_nhwc(a, dt, cuda), _nhwc(b_, dt, cuda), wide[..., 64:]], _pack(w, dt, cuda), outbuf[..., 64:], kh=3, kw=3, pad=1,
                  bias=bias.to(cuda))
    _close(outbuf[..., 64:].permute(0, 3, 1, 2), ref, dt)
    assert outbuf[..., :64].abs().max().item() == 0


@pytest.mark.parametrize("dt", DTYPES)
def test_conv4x4_stride2(cuda, dt):
    from fbanet_b200 import ops
    x, w, b = _r(dt, 2, 64, 20, 20, seed=1), _r(dt, 128, 64, 4, 4, seed=2, scale=0.03), _r(torch.float32, 128, seed=3)
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    out = torch.empty(2, 10, 10, 128, device=cuda, dtype=dt)
    ops.conv_gemm([_nhwc(x, dt, cuda)], _pack(w, dt, cuda), out, kh=4, kw=4, stride=2, pad=1, bias=b.to(cuda))
    _close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_transpose2x2(cuda, dt):
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(dt, 2, 128, 10, 10, seed=1), _r(dt, 128, 64, 2, 2, seed=2, scale=0.05), _r(torch.float32, 64, seed=3)
    ref = F.conv_transpose2d(x, w, b, stride=2)
    wp = w.permute(2, 3, 1, 0).reshape(4 * 64, 128).contiguous().to(cuda, dt)
    out = torch.empty(2, 20, 20, 64, device=cuda, dtype=dt)
    ops.conv_gemm([_nhwc(x, dt, cuda)], wp, out, bias=b.repeat(4).to(cuda), store_mode=L.STORE_CONVT2)
    _close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_pixelshuffle_store(cuda, dt):
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(dt, 1, 64, 12, 12, seed=1), _r(dt, 256, 64, 3, 3, seed=2, scale=0.03), _r(torch.float32, 256, seed=3)
    ref = F.pixel_shuffle(F.conv2d(x, w, b, padding=1), 2)
    out = torch.empty(1, 24, 24, 64, device=cuda, dtype=dt)
    ops.conv_gemm([_nhwc(x, dt, cuda)], _pack(w, dt, cuda), out, kh=3, kw=3, pad=1, bias=b.to(cuda), store_mode=L.STORE_PS2)
    _close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("dt", DTYPES)
def test_final_conv_with_bilinear_base(cuda, dt):
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(dt, 2, 64, 32, 32, seed=1), _r(dt, 3, 64, 3, 3, seed=2, scale=0.03), _r(torch.float32, 3, seed=3)
    burst = torch.rand(2, 5, 3, 8, 8, generator=torch.Generator().manual_seed(4))
    ref = F.conv2d(x, w, b, padding=1) + F.interpolate(burst[:, 0], scale_factor=4, mode="bilinear", align_corners=False)
    bd = burst.to(cuda)
    out = torch.empty(2, 3, 32, 32, device=cuda, dtype=torch.float32)
    ops.conv_gemm([_nhwc(x, dt, cuda)], _pack(w, dt, cuda), out, kh=3, kw=3, pad=1, bias=b.to(cuda), store_mode=L.STORE_NCHW_BASE, base=bd[:, 0])
    _close(out, ref, dt)


@pytest.mark.parametrize("dt", DTYPES)
def test_linear_gelu_residual(cuda, dt):
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(dt, 1, 300, 128, seed=1), _r(dt, 512, 128, seed=2, scale=0.08), _r(torch.float32, 512, seed=3)
    ref = F.gelu(F.linear(x, w, b), approximate="tanh")
    out = torch.empty(1, 15, 20, 512, device=cuda, dtype=dt)
    ops.conv_gemm([x.view(1, 15, 20, 128).to(cuda, dt)], w.to(cuda, dt), out, bias=b.to(cuda), act=L.ACT_GELU_TANH)
    _close(out.view(1, 300, 512), ref, dt)


@pytest.mark.parametrize("dt", DTYPES)
def test_faf_fuse_1x1_with_gate(cuda, dt):
    """K2: 14-source 1x1 conv with the per-pixel gate applied to frames 1..F-1 (federated_affinity_fusion.py:95-128)."""
    from fbanet_b200 import ops, _lib as L
    B, Fr, E, S = 2, 14, 64, 12
    feat = _r(dt, B, Fr, E, S, S, seed=1)
    gate = torch.rand(B, Fr - 1, S, S, generator=torch.Generator().manual_seed(2)) * 0.5 + 0.5
    w, b = _r(dt, E, Fr * E, 1, 1, seed=3, scale=0.03), _r(torch.float32, E, seed=4)
    g = torch.cat([feat[:, :1], feat[:, 1:] * gate[:, :, None]], 1)
    ref = F.prelu(F.conv2d(g.reshape(B, Fr * E, S, S), w, b), torch.tensor([0.1]))
    fd = feat.permute(0, 1, 3, 4, 2).contiguous().to(cuda, dt)
    gd = gate.to(cuda)
    out = torch.empty(B, S, S, E, device=cuda, dtype=dt)
    ops.conv_gemm([fd[:, f] for f in range(Fr)], w.reshape(E, -1).to(cuda, dt), out, bias=b.to(cuda), act=L.ACT_PRELU,
                  alpha=torch.tensor([0.1], device=cuda), row_scales=[None] + [gd[:, f - 1] for f in range(1, Fr)])
    _close(out.permute(0, 3, 1, 2), ref, dt, scale=2.0)


@pytest.mark.parametrize("dt", DTYPES)
def test_faf_gate_matches_reference_formula(cuda, dt):
    """gate kernel (collapsed form) vs the reference's as-written affinity maths in float64."""
    from fbanet_b200 import ops
    B, Fr, E, S = 2, 6, 64, 14
    feat = _r(dt, B, Fr, E, S, S, seed=1)
    w0, b0 = _r(torch.float32, E, E, 3, 3, seed=2, scale=0.04), _r(torch.float32, E, seed=3)
    w1, b1 = _r(torch.float32, E, E, 3, 3, seed=4, scale=0.04), _r(torch.float32, E, seed=5)
    fd64 = feat.double()
    ref_e = F.conv2d(fd64[:, 0], w0.double(), b0.double(), padding=1)
    emb = F.conv2d(fd64.reshape(B * Fr, E, S, S), w1.double(), b1.double(), padding=1).view(B, Fr, E, S, S)
    aff = (emb - ref_e[:, None]).sum(2)
    ref = torch.sigmoid((aff[:, 1:] - aff[:, :1]).abs())
    wsum = w1.double().sum(0).permute(1, 2, 0).reshape(9, E).float().contiguous().to(cuda)
    gate = ops.faf_gate(feat.permute(0, 1, 3, 4, 2).contiguous().to(cuda, dt), wsum)
    assert (gate.cpu().double() - ref).abs().max().item() < 2e-5


def test_faf_gate_tensor_core_scores(cuda):
    """bf16 path: the wsum dot products as a 3x3 implicit GEMM (hi/lo weight rows, fp32 NHWC_F32 store) feeding the streaming
    gate-apply kernel == the as-written affinity maths in float64 (blocks/federated_affinity_fusion.py:79-105), and the
    gated features == feat * gate in the pixel-major layout of the 1x1 fusion GEMM."""
    from fbanet_b200 import ops
    dt = torch.bfloat16
    B, Fr, E, S = 2, 6, 64, 24
    feat = _r(dt, B, Fr, E, S, S, seed=1)
    w1 = _r(torch.float32, E, E, 3, 3, seed=4, scale=0.04)
    fd64 = feat.double()
    emb = F.conv2d(fd64.reshape(B * Fr, E, S, S), w1.double(), None, padding=1).view(B, Fr, E, S, S)
    aff = emb.sum(2)
    ref = torch.sigmoid((aff[:, 1:] - aff[:, :1]).abs())
    wsum = w1.double().sum(0).permute(1, 2, 0).reshape(9, E).float().contiguous().to(cuda)
    fd = feat.permute(0, 1, 3, 4, 2).contiguous().to(cuda, dt)
    score = ops.faf_scores(fd, ops.faf_score_weight(wsum, dt))
    s_ref = aff.reshape(B * Fr, S, S)
    assert ((score[..., 0] + score[..., 1]).cpu().double() - s_ref).abs().max().item() < 2e-4 * max(1.0, s_ref.abs().max().item())
    gate, gated = ops.faf_gate(fd, wsum, want_gate=True, want_gated=True, score=score)
    assert (gate.cpu().double() - ref).abs().max().item() < 1e-4
    g_all = torch.cat([torch.ones(B, 1, S, S, dtype=torch.float64), ref], 1)                      # frame 0 passes through
    want = (fd64 * g_all[:, :, None]).permute(0, 3, 4, 1, 2).reshape(B, S, S, Fr * E)             # [B,H,W,F*C]
    _close(gated, want.float(), dt)
    # and the same gate as the CUDA-core kernel computes on its own
    gate2 = ops.faf_gate(fd, wsum)
    assert (gate2 - gate).abs().max().item() < 1e-4


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("C", [64, 128, 256])
def test_layernorm(cuda, dt, C):
    from fbanet_b200 import ops
    x = _r(dt, 1000, C, seed=1, scale=2.0) + 0.3
    g, b = _r(torch.float32, C, seed=2) + 1.0, _r(torch.float32, C, seed=3)
    ref = F.layer_norm(x, (C,), g, b, 1e-5)
    y = ops.layernorm(x.to(cuda, dt), g.to(cuda), b.to(cuda))
    _close(y, ref, dt, scale=2.0)


@pytest.mark.parametrize("C,cout,act", [(64, 192, 0), (128, 512, 3), (256, 768, 0), (128, 384, 0)])
def test_layernorm_folded_into_gemm(cuda, C, cout, act):
    """bf16 path: row statistics + GEMM on the raw rows with gamma folded into the weights and the (mean, rstd) correction in
    the epilogue (row-centred weights, rstd scaling) == Linear(LayerNorm(x)) (layers/fba_net.py:196,246 -> linear_projection.py:27-28 /
    locally_enhanced_feed_forward.py:27), including rows with a large mean."""
    from fbanet_b200 import ops, _lib as L
    dt = torch.bfloat16
    N, H, W = 3, 24, 16
    x = _r(dt, N, H, W, C, seed=1, scale=2.0) + _r(dt, N, H, W, 1, seed=5, scale=30.0)   # per-row offsets: |mean| up to 25 sigma
    x = x.to(dt).float()
    g, be = _r(torch.float32, C, seed=2) + 1.0, _r(torch.float32, C, seed=3)
    w, b = _r(torch.float32, cout, C, seed=4, scale=1 / math.sqrt(C)), _r(torch.float32, cout, seed=6, scale=0.1)
    ref = F.linear(F.layer_norm(x.double(), (C,), g.double(), be.double(), 1e-5), w.double(), b.double()).float()
    if act == 3:
        ref = F.gelu(ref, approximate="tanh")
    xd = x.to(cuda, dt)
    st = ops.row_stats(xd.view(-1, C))
    mean, var = x.double().mean(-1), x.double().var(-1, unbiased=False)
    assert (st[:, 0].cpu().double() - mean.view(-1)).abs().max().item() < 1e-5
    assert ((st[:, 1].cpu().double() - (var.view(-1) + 1e-5).rsqrt()).abs() / (var.view(-1) + 1e-5).rsqrt()).max().item() < 1e-5
    wf, bf = ops.fold_layernorm(w.to(cuda), b.to(cuda), g.to(cuda), be.to(cuda), dt)
    assert wf.double().sum(1).abs().max().item() < 1e-4   # rows centred to ~2^-18: the mean subtraction lives in the weights
    out = torch.empty(N, H, W, cout, device=cuda, dtype=dt)
    ops.conv_gemm([xd], wf, out, bias=bf, act=act, ln_stats=st, impl=L.IMPL_TCGEN05)
    _close(out, ref, dt, scale=2.0)


@pytest.mark.parametrize("dt", DTYPES)
def test_dwconv_gelu(cuda, dt):
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(dt, 2, 256, 14, 10, seed=1), _r(torch.float32, 256, 1, 3, 3, seed=2, scale=0.3), _r(torch.float32, 256, seed=3)
    ref = F.gelu(F.conv2d(x, w, b, padding=1, groups=256), approximate="tanh")
    y = ops.dwconv3x3(_nhwc(x, dt, cuda), w.reshape(256, 9).t().contiguous().to(cuda), b.to(cuda), L.ACT_GELU_TANH)
    _close(y.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("C,heads,HW,win,shift", [(64, 1, 20, 10, 0), (64, 1, 20, 10, 5), (128, 2, 30, 10, 5), (256, 16, 20, 10, 5),
                                                   (128, 8, 20, 10, 0), (64, 4, 5, 5, 0), (128, 8, 8, 4, 2), (128, 8, 40, 10, 5),
                                                   (256, 16, 30, 10, 3)])
@pytest.mark.parametrize("prescaled", [False, True])
def test_window_attention(cuda, dt, C, heads, HW, win, shift, prescaled):
    """K6 against the oracle's WindowAttention + roll/partition/mask (layers/fba_net.py:139-250)."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import shift_attn_mask, window_partition, window_reverse, relative_position_index
    if prescaled and dt != torch.bfloat16:
        pytest.skip("q_prescaled is a tensor-core (bf16) path option")
    B, H, W = 2, HW, HW
    N, dh = win * win, C // heads
    qkv = _r(dt, B, H, W, 3 * C, seed=1, scale=1.5)
    table = _r(torch.float32, (2 * win - 1) ** 2, heads, seed=2, scale=0.5)
    scale = dh ** -0.5
    y = qkv
    if shift:
        y = torch.roll(y, (-shift, -shift), (1, 2))
    yw = window_partition(y, win)  # [B*nW, N, 3C]
    q, k, v = (yw[..., i * C:(i + 1) * C].view(-1, N, heads, dh).permute(0, 2, 1, 3) for i in range(3))
    attn = (q * scale) @ k.transpose(-2, -1)
    bias = table[relative_position_index(win).view(-1)].view(N, N, heads).permute(2, 0, 1)
    attn = attn + bias[None]
    if shift:
        mask = shift_attn_mask(H, W, win, shift)
        nW = mask.shape[0]
        attn = (attn.view(B, nW, heads, N, N) + mask[None, :, None]).view(-1, heads, N, N)
    o = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(-1, N, C)
    o = window_reverse(o, win, B, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    qkv_in = qkv.reshape(-1, 3 * C).clone()
    if prescaled:  # the caller folds scale*log2(e) into the q projection (model.py:_pack)
        qkv_in[:, :C] = qkv_in[:, :C] * (scale * 1.4426950408889634)
    got = ops.window_attention(qkv_in.to(cuda, dt), table.to(cuda), B, H, W, heads, win, shift, scale, q_prescaled=prescaled)
    _close(got.view(B, H, W, C), o, dt, scale=2.0)


@pytest.mark.parametrize("C,heads,HW,shift,B,mag", [(64, 1, 20, 0, 2, 1.5), (64, 1, 20, 5, 2, 1.5), (128, 2, 30, 5, 2, 1.5), (64, 1, 40, 5, 1, 1.5),
                                                     (128, 2, 40, 0, 3, 1.5), (64, 1, 160, 5, 2, 1.5), (128, 2, 30, 5, 1, 10.0)])
def test_window_attention_tcgen05_dh64(cuda, C, heads, HW, shift, B, mag):
    """K6 for the d_h = 64 stages on tcgen05 + TMEM (S = Q K^T and O = P V as tcgen05.mma, scores and P in tensor memory / shared
    memory, one softmax thread per query row) against the oracle's roll + partition + WindowAttention + mask
    (layers/fba_net.py:139-250, layers/window_attention.py:173-243): unshifted and shifted layers -- with 2, 3, 4 and 16 windows
    per side every wrap type occurs (none, bottom edge, right edge, corner: the permuted row order and the block mask) -- one and
    two heads, more items than SMs, and logits in the hundreds (exact row max)."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import shift_attn_mask, window_partition, window_reverse, relative_position_index
    dt, win = torch.bfloat16, 10
    H = W = HW
    N, dh = win * win, C // heads
    scale = dh ** -0.5
    qkv = _r(dt, B, H, W, 3 * C, seed=1, scale=mag)
    qkv[..., 2 * C:] = _r(dt, B, H, W, C, seed=6, scale=1.0)
    table = _r(torch.float32, (2 * win - 1) ** 2, heads, seed=2, scale=0.5)
    qkv_in = qkv.reshape(-1, 3 * C).clone()
    qkv_in[:, :C] = (qkv_in[:, :C] * (scale * 1.4426950408889634)).to(dt).float()      # what the folded q projection emits
    qkv = qkv.clone()
    qkv[..., :C] = qkv_in[:, :C].view(B, H, W, C) / (scale * 1.4426950408889634)       # the reference sees the same rounded q
    y = torch.roll(qkv, (-shift, -shift), (1, 2)) if shift else qkv
    yw = window_partition(y, win)
    q, k, v = (yw[..., i * C:(i + 1) * C].view(-1, N, heads, dh).permute(0, 2, 1, 3) for i in range(3))
    attn = (q.double() * scale) @ k.double().transpose(-2, -1)
    attn = attn + table[relative_position_index(win).view(-1)].view(N, N, heads).permute(2, 0, 1)[None].double()
    if shift:
        mask = shift_attn_mask(H, W, win, shift)
        attn = (attn.view(B, mask.shape[0], heads, N, N) + mask[None, :, None].double()).view(-1, heads, N, N)
    o = (torch.softmax(attn, -1) @ v.double()).float().transpose(1, 2).reshape(-1, N, C)
    o = window_reverse(o, win, B, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    tab = table.to(cuda)
    bx = ops.expand_rel_pos_bias(tab, win)
    got = ops.window_attention(qkv_in.to(cuda, dt), tab, B, H, W, heads, win, shift, scale, bias_expanded=bx, q_prescaled=True,
                               bias_wrap=ops.expand_rel_pos_bias_wrap(bx, win) if shift else None)
    assert ops.LAST_ATTENTION_ON_TCGEN05, "the d_h = 64 shapes must take the tcgen05 kernel"
    assert torch.isfinite(got.float()).all()
    _close(got.view(B, H, W, C), o, dt, scale=2.0)


@pytest.mark.parametrize("prescaled", [False, True])
@pytest.mark.parametrize("mag", [6.0, 14.0])
def test_window_attention_huge_logits(cuda, prescaled, mag):
    """dh=16 kernel: logits far outside exp2's range must take the exact-row-max path (and moderate ones the shift-free fast
    path) and still match softmax with max subtraction (layers/window_attention.py:230)."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import shift_attn_mask, window_partition, window_reverse, relative_position_index
    dt, B, H, W, C, heads, win, shift = torch.bfloat16, 1, 30, 30, 128, 8, 10, 5
    N, dh = win * win, C // heads
    qkv = _r(dt, B, H, W, 3 * C, seed=5, scale=mag)
    qkv[..., 2 * C:] = _r(dt, B, H, W, C, seed=6, scale=1.0)
    table = _r(torch.float32, (2 * win - 1) ** 2, heads, seed=2, scale=0.5)
    scale = dh ** -0.5
    qkv_in = qkv.reshape(-1, 3 * C).clone()
    if prescaled:
        qkv_in[:, :C] = (qkv_in[:, :C] * (scale * 1.4426950408889634)).to(dt).float()
        qkv = qkv.clone()
        qkv[..., :C] = qkv_in[:, :C].view(B, H, W, C) / (scale * 1.4426950408889634)  # reference sees the same rounded q
    yw = window_partition(torch.roll(qkv, (-shift, -shift), (1, 2)), win)
    q, k, v = (yw[..., i * C:(i + 1) * C].view(-1, N, heads, dh).permute(0, 2, 1, 3) for i in range(3))
    attn = (q.double() * scale) @ k.double().transpose(-2, -1)
    attn = attn + table[relative_position_index(win).view(-1)].view(N, N, heads).permute(2, 0, 1)[None].double()
    mask = shift_attn_mask(H, W, win, shift)
    attn = (attn.view(B, mask.shape[0], heads, N, N) + mask[None, :, None].double()).view(-1, heads, N, N)
    assert attn.abs().max().item() > (300.0 if mag > 10 else 50.0)
    o = (torch.softmax(attn, -1) @ v.double()).float().transpose(1, 2).reshape(-1, N, C)
    o = torch.roll(window_reverse(o, win, B, H, W), (shift, shift), (1, 2))
    got = ops.window_attention(qkv_in.to(cuda, dt), table.to(cuda), B, H, W, heads, win, shift, scale, q_prescaled=prescaled)
    assert torch.isfinite(got.float()).all()
    _close(got.view(B, H, W, C), o, dt, scale=2.0)


def test_to_nhwc(cuda):
    from fbanet_b200 import ops
    x = torch.rand(6, 3, 9, 7, generator=torch.Generator().manual_seed(0))
    for dt, cp in ((torch.float32, 4), (torch.bfloat16, 8)):
        y = ops.to_nhwc(x.to(cuda), cp, dt).cpu().float()
        assert torch.equal(y[..., :3], x.permute(0, 2, 3, 1).to(dt).float())
        assert y[..., 3:].abs().max().item() == 0


def _rand_homographies(B, T, seed=1):
    """H = I + eps; translation U(-4,4), affine U(-.01,.01), perspective U(-1e-5,1e-5)."""
    g = np.random.default_rng(seed)
    M = np.tile(np.eye(3), (B, T, 1, 1))
    M[..., :2, :2] += g.uniform(-0.01, 0.01, (B, T, 2, 2))
    M[..., :2, 2] += g.uniform(-4, 4, (B, T, 2))
    M[..., 2, :2] += g.uniform(-1e-5, 1e-5, (B, T, 2))
    M[:, 0] = np.eye(3)
    return M


@pytest.mark.parametrize("layout", ["BTCHW", "BTHWC"])
def test_warp_matches_float64_oracle(cuda, layout):
    """K1: coordinates within 1e-5 px of the float64 closed form, samples to fp32 rounding."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import warp_frame, warp_coords
    B, T, Cc, H, W = 2, 5, 4, 40, 56
    burst = torch.rand(B, T, H, W, Cc, generator=torch.Generator().manual_seed(0))
    M = _rand_homographies(B, T)
    src = burst if layout == "BTHWC" else burst.permute(0, 1, 4, 2, 3).contiguous()
    out, coords = ops.warp_burst(src.to(cuda), torch.from_numpy(M), layout=layout, return_coords=True)
    out = out.cpu() if layout == "BTHWC" else out.cpu().permute(0, 1, 3, 4, 2)
    for b in range(B):
        assert torch.equal(out[b, 0], burst[b, 0])  # base frame untouched
        for t in range(1, T):
            sx, sy = warp_coords(M[b, t], H, W)
            c = coords[b, t].cpu().numpy()
            assert np.abs(c[..., 0] - sx).max() < 1e-5 and np.abs(c[..., 1] - sy).max() < 1e-5
            ref = warp_frame(burst[b, t].numpy(), M[b, t])
            assert np.abs(out[b, t].numpy() - ref).max() < 2e-6


@pytest.mark.parametrize("shape", [(3, 48, 160), (4, 80, 80), (3, 20, 36), (4, 33, 64)])
@pytest.mark.parametrize("kind", ["small", "zoom", "rotate", "horizon"])
def test_warp_tile_staged_kernel_and_its_fallbacks(cuda, shape, kind):
    """The planar warp kernel stages the source box of every 32 x 16 (or 16 x 32) destination tile in shared memory; tiles whose box
    does not fit (zoom, rotation) or whose w changes sign (a horizon inside the frame) gather from global memory instead.  Every
    family against the float64 oracle, on widths that use either tile shape, ragged edge tiles included."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import warp_frame
    Cc, H, W = shape
    B, T = 2, 4
    g = np.random.default_rng(11)
    burst = torch.rand(B, T, Cc, H, W, generator=torch.Generator().manual_seed(3))
    M = np.tile(np.eye(3), (B, T, 1, 1))
    for b in range(B):
        for t in range(1, T):
            if kind == "small":
                M[b, t, :2, :2] += g.uniform(-0.01, 0.01, (2, 2)); M[b, t, :2, 2] = g.uniform(-4, 4, 2); M[b, t, 2, :2] = g.uniform(-1e-5, 1e-5, 2)
            elif kind == "zoom":
                M[b, t, 0, 0] = M[b, t, 1, 1] = g.uniform(1.5, 2.5); M[b, t, :2, 2] = g.uniform(-20, 5, 2)
            elif kind == "rotate":
                a = g.uniform(0.5, 1.2); M[b, t, :2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]; M[b, t, :2, 2] = g.uniform(0, W / 2, 2)
            else:   # w = 1 - x / (W/2): changes sign in the middle of the frame
                M[b, t, 2, 0] = -2.0 / W; M[b, t, :2, 2] = g.uniform(-3, 3, 2)
    out = ops.warp_burst(burst.to(cuda), torch.from_numpy(M)).cpu()
    for b in range(B):
        assert torch.equal(out[b, 0], burst[b, 0])
        for t in range(1, T):
            ref = warp_frame(burst[b, t].permute(1, 2, 0).numpy(), M[b, t])
            got = out[b, t].permute(1, 2, 0).numpy()
            if kind == "horizon":   # next to the horizon the coordinates are huge and ill conditioned in any arithmetic: compare away from it
                keep = np.abs(1.0 - 2.0 * np.arange(W) / W) > 0.05
                ref, got = ref[:, keep], got[:, keep]
            assert np.abs(got - ref).max() < 2e-6, (b, t, np.abs(got - ref).max())


def test_warp_large_frame_coordinates(cuda):
    """fp32 cannot represent 1e-5 px at x ~ 1900; the kernel evaluates coordinates in fp64."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import warp_coords
    H, W = 270, 1920
    burst = torch.rand(1, 2, 1, H, W, generator=torch.Generator().manual_seed(0))
    M = _rand_homographies(1, 2, seed=3)
    _, coords = ops.warp_burst(burst.to(cuda), torch.from_numpy(M), return_coords=True)
    sx, sy = warp_coords(M[0, 1], H, W)
    c = coords[0, 1].cpu().numpy()
    assert np.abs(c[..., 0] - sx).max() < 1e-5 and np.abs(c[..., 1] - sy).max() < 1e-5


def test_warp_vs_cv2_quantised(cuda):
    """Secondary pin: cv2.warpPerspective == bilinear with coords rounded to 1/32 px; our exact
    warp must agree with cv2 to the size of that quantisation on a smooth image."""
    cv2 = pytest.importorskip("cv2")
    from fbanet_b200 import ops
    H, W = 48, 64
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    img = np.stack([np.sin(xx / 9.0) * np.cos(yy / 7.0), xx / W, yy / H], -1).astype(np.float32)
    M = _rand_homographies(1, 2, seed=5)
    ref = cv2.warpPerspective(img, M[0, 1], (W, H), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)
    burst = torch.from_numpy(np.stack([img, img])[None])
    out = ops.warp_burst(burst.to(cuda), torch.from_numpy(M), layout="BTHWC").cpu().numpy()[0, 1]
    inner = (slice(6, -6), slice(6, -6))
    assert np.abs(out[inner] - ref[inner]).max() < 5e-3


def test_tile_divide_merge(cuda):
    """8f-1: GPU reflect-pad tiling + stitch vs the restated utils/dataset_utils.py."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import tensor_divide_burst, tensor_merge
    T, Cc, H, W, ps, ov = 3, 3, 50, 70, 20, 10
    burst = torch.rand(1, T, Cc, H, W, generator=torch.Generator().manual_seed(0))
    ref = tensor_divide_burst(burst, ps, ov)
    got = ops.tile_divide(burst[0].to(cuda), ps, ov)
    assert torch.equal(got.cpu(), ref)
    sr = torch.rand(ref.shape[0], Cc, 4 * (ps + 2 * ov), 4 * (ps + 2 * ov), generator=torch.Generator().manual_seed(1))
    refm = tensor_merge(sr, (4 * H, 4 * W), 4 * ps, 4 * ov)
    out = torch.zeros(Cc, 4 * H, 4 * W, device=cuda)
    ops.tile_merge(sr.to(cuda), out, H, W, ps, ov, 4)
    assert torch.equal(out.cpu(), refm[0])


@pytest.mark.parametrize("nbands", [1, 2, 3, 8])
def test_tile_divide_merge_banded(cuda, nbands):
    """cfg4 / 8e: the row-band kernels (bands possibly in peer memory; here all on one GPU, as separate allocations) must gather
    exactly the tiles, and stitch exactly the image, of the restated utils/dataset_utils.py -- including tiles whose halo or
    core crosses band boundaries and bands thinner than the halo."""
    from fbanet_b200 import ops
    from fbanet_b200.dist import band_rows, shard_range
    from oracle.fbanet_oracle import tensor_divide_burst, tensor_merge
    T, Cc, H, W, ps, ov = 3, 3, 50, 70, 20, 10
    burst = torch.rand(1, T, Cc, H, W, generator=torch.Generator().manual_seed(0))
    ref = tensor_divide_burst(burst, ps, ov)
    row0 = band_rows(H, nbands)
    bands = [burst[0, :, :, row0[k]:row0[k + 1]].contiguous().to(cuda) for k in range(nbands)]
    ntiles = ref.shape[0]
    sr = torch.rand(ntiles, Cc, 4 * (ps + 2 * ov), 4 * (ps + 2 * ov), generator=torch.Generator().manual_seed(1))
    refm = tensor_merge(sr, (4 * H, 4 * W), 4 * ps, 4 * ov)[0]
    obands = [torch.full((Cc, 4 * (row0[k + 1] - row0[k]), 4 * W), -1.0, device=cuda) for k in range(nbands)]
    for r in range(2):                                  # two tile shards, as two ranks would run them
        b, e = shard_range(ntiles, r, 2)
        got = ops.tile_divide_banded([x.data_ptr() for x in bands], row0, T, Cc, H, W, ps, ov, b, e, cuda)
        assert torch.equal(got.cpu(), ref[b:e])
        ops.tile_merge_banded(sr[b:e].contiguous().to(cuda), [x.data_ptr() for x in obands], row0, H, W, ps, ov, 4, b, e)
    assert torch.equal(torch.cat([x.cpu() for x in obands], 1), refm)


@pytest.mark.parametrize("layout", ["BTCHW", "BTHWC"])
@pytest.mark.parametrize("shape", [(2, 5, 3, 37, 53), (1, 14, 4, 80, 80), (3, 1, 3, 16, 24), (1, 3, 2, 9, 7)])
def test_flow_warp_matches_oracle(cuda, layout, shape):
    """8f-4: registration/optical_flow/register.py:11-47 -- bilinear sample at grid - flow with edge-clamped indices.  Flows reach
    far outside the image, include exact integers and NaN-free extremes; frame 0 has no flow and is copied."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import flow_register_burst, flow_register_frame
    B, T, Cc, H, W = shape
    g = torch.Generator().manual_seed(3)
    burst = torch.rand(B, T, H, W, Cc, generator=g)
    nf = 1 if T == 1 else T - 1
    flow = (torch.rand(B, nf, H, W, 2, generator=g) - 0.5) * 2 * max(H, W) * 0.6
    flow[:, :, ::3, ::2] = torch.round(flow[:, :, ::3, ::2])          # exact-integer displacements
    flow[:, :, 0, 0] = torch.tensor([1e9, -1e9])                      # wild values must clamp, not overflow
    if T == 1:
        ref = np.stack([flow_register_frame(burst[b, 0].numpy(), flow[b, 0].numpy())[None] for b in range(B)])
    else:
        ref = np.stack([flow_register_burst(burst[b].numpy(), flow[b].numpy()) for b in range(B)])
    x = burst if layout == "BTHWC" else burst.permute(0, 1, 4, 2, 3).contiguous()
    out = ops.flow_warp_burst(x.to(cuda), flow.to(cuda), layout=layout).cpu()
    out = out if layout == "BTHWC" else out.permute(0, 1, 3, 4, 2)
    assert np.abs(out.numpy() - ref).max() <= 2e-6, np.abs(out.numpy() - ref).max()
    if T > 1:
        assert torch.equal(out[:, 0], burst[:, 0])


def test_flow_warp_zero_flow_is_identity_and_shift_is_exact(cuda):
    from fbanet_b200 import ops
    x = torch.rand(1, 3, 3, 32, 40, generator=torch.Generator().manual_seed(0)).to(cuda)
    assert torch.equal(ops.flow_warp_burst(x, torch.zeros(1, 2, 32, 40, 2)), x)
    fl = torch.zeros(1, 2, 32, 40, 2)
    fl[..., 0], fl[..., 1] = 2.0, -3.0                                # out(y,x) = in(y-2, x+3), clamped at the edges
    out = ops.flow_warp_burst(x, fl)
    yy = (torch.arange(32) - 2).clamp(0, 31)
    xx = (torch.arange(40) + 3).clamp(0, 39)
    assert torch.equal(out[:, 1:].cpu(), x.cpu()[:, 1:][..., yy, :][..., xx])


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("cin,hw", [(3, (21, 33)), (4, (16, 16))])
def test_head_conv_direct(cuda, dt, cin, hw):
    """models/fba_net.py:255 head conv straight from the planar burst (CUDA-core kernel)."""
    from fbanet_b200 import ops
    H, W = hw
    x = torch.rand(5, cin, H, W, generator=torch.Generator().manual_seed(1))
    w, b = _r(torch.float32, 64, cin, 3, 3, seed=2, scale=0.2), _r(torch.float32, 64, seed=3)
    ref = F.conv2d(x, w, b, padding=1)
    wkc = w.permute(2, 3, 1, 0).reshape(9 * cin, 64).contiguous()
    out = ops.head_conv(x.to(cuda), wkc.to(cuda), b.to(cuda), dt)
    _close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("cin,hw,frames", [(3, (32, 24), 5), (4, (20, 44), 3), (3, (160, 160), 14)])
def test_head_conv_tensor_core_keeps_fp32_samples(cuda, cin, hw, frames):
    """bf16 head conv on tcgen05: samples enter as hi + lo bf16 halves, so only the weights are rounded to bf16 -- the
    result must match the fp32 conv with bf16-rounded weights to bf16 output rounding, including ragged tiles."""
    from fbanet_b200 import ops
    H, W = hw
    x = torch.rand(frames, cin, H, W, generator=torch.Generator().manual_seed(1))
    w, b = _r(torch.float32, 64, cin, 3, 3, seed=2, scale=0.2), _r(torch.float32, 64, seed=3)
    ref = F.conv2d(x.double(), w.bfloat16().double(), b.double(), padding=1).float()
    wkc = w.permute(2, 3, 1, 0).reshape(9 * cin, 64).contiguous()
    out = ops.head_conv(x.to(cuda), wkc.to(cuda), b.to(cuda), torch.bfloat16).float().cpu().permute(0, 3, 1, 2)
    err = (out - ref).abs()
    assert (err <= 2.0 ** -8 * ref.abs() + 1e-4).all(), err.max().item()   # bf16 rounding of the output only


@pytest.mark.parametrize("dt", DTYPES)
def test_assemble_sr_plus_bilinear_base(cuda, dt):
    """models/fba_net.py:317-320: channels-last SR + bilinear x4 (align_corners=False) of frame 0 -> planar fp32."""
    from fbanet_b200 import ops
    sr = _r(dt, 2, 32, 48, 8, seed=1)
    burst = torch.rand(2, 5, 3, 8, 12, generator=torch.Generator().manual_seed(2))
    ref = sr[..., :3].permute(0, 3, 1, 2) + F.interpolate(burst[:, 0], scale_factor=4, mode="bilinear", align_corners=False)
    out = ops.assemble(sr.to(cuda, dt), burst.to(cuda)[:, 0], 3)
    assert (out.cpu() - ref).abs().max().item() < 1e-5


def test_ecc_prepare_matches_oracle(cuda):
    """8f-4: grey conversion + 5x5 Gaussian + central differences of fbanet_ecc_prepare_sm100 vs the cv2-pinned restatement."""
    from fbanet_b200 import _lib as L, ops  # noqa: F401
    from oracle.fbanet_oracle import bgr2gray, ecc_blur5, ecc_gradients
    import ctypes as C
    g = torch.Generator().manual_seed(5)
    burst = torch.rand(2, 3, 3, 37, 50, generator=g)
    x = burst.to(cuda)
    planes = torch.empty((6, 3, 37, 50), device=cuda)
    pp = L.EccPrepareParams()
    pp.src, pp.planes = x.data_ptr(), planes.data_ptr()
    pp.s_frame, pp.s_c, pp.s_y, pp.s_x = 3 * 37 * 50, 37 * 50, 50, 1
    for c, w in enumerate((0.114, 0.587, 0.299, 0.0)):
        pp.gray_weight[c] = w
    pp.frames, pp.H, pp.W, pp.C = 6, 37, 50, 3
    ops._call("fbanet_ecc_prepare_sm100", pp)
    got = planes.cpu().numpy()
    for f in range(6):
        b = ecc_blur5(bgr2gray(burst.view(6, 3, 37, 50)[f].permute(1, 2, 0).numpy()))
        gx, gy = ecc_gradients(b)
        assert np.abs(got[f, 0] - b).max() < 1e-6 and np.abs(got[f, 1] - gx).max() < 1e-6 and np.abs(got[f, 2] - gy).max() < 1e-6


@pytest.mark.parametrize("layout", ["BTCHW", "BTHWC"])
def test_ecc_homography_matches_oracle_and_registers(cuda, layout):
    """8f-4: the on-device ECC iterations (fp32 pixel maths, fp64 reductions and solve) against the float64 restatement of
    cv2.findTransformECC, on bursts whose frames are known homographies of the base frame + noise; then the estimated matrices,
    fed to the K1 warp, must register the frames (PSNR to the base frame goes up)."""
    import cv2
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import bgr2gray, ecc_homography, homography_coord_diff as _coord_diff
    H, W, T, B = 96, 128, 4, 2
    rng = np.random.default_rng(7)
    burst = np.zeros((B, T, H, W, 3), np.float32)
    for b in range(B):
        base = cv2.GaussianBlur(rng.random((H, W, 3)).astype(np.float32), (0, 0), 2.5)
        base = (base - base.min()) / (base.max() - base.min())
        burst[b, 0] = base
        for t in range(1, T):
            M = np.eye(3, dtype=np.float32)
            M[0, 2], M[1, 2] = rng.uniform(-3, 3), rng.uniform(-3, 3)
            M[0, 1], M[1, 0] = rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01)
            M[2, 1] = rng.uniform(-1e-5, 1e-5)
            burst[b, t] = cv2.warpPerspective(base, M, (W, H), flags=cv2.INTER_LINEAR) + rng.normal(0, 0.01, (H, W, 3)).astype(np.float32)
    xb = torch.from_numpy(burst)
    x = (xb if layout == "BTHWC" else xb.permute(0, 1, 4, 2, 3).contiguous()).to(cuda)
    M, rho, iters = ops.ecc_homography_burst(x, layout=layout)
    M, rho, iters = M.cpu().numpy(), rho.cpu().numpy(), iters.cpu().numpy()
    assert np.array_equal(M[:, 0], np.tile(np.eye(3), (B, 1, 1))) and (iters[:, 1:] > 0).all()
    for b in range(B):
        g0 = bgr2gray(burst[b, 0])
        for t in range(1, T):
            r_ref, M_ref = ecc_homography(g0, bgr2gray(burst[b, t]))
            assert abs(rho[b, t] - r_ref) < 2e-4, (rho[b, t], r_ref)
            assert _coord_diff(M[b, t], M_ref, H, W) < 5e-3, _coord_diff(M[b, t], M_ref, H, W)
    reg = ops.warp_burst(x, torch.from_numpy(M), layout=layout).cpu()
    reg = reg if layout == "BTHWC" else reg.permute(0, 1, 3, 4, 2)

    def psnr_in(a, c):   # inner region: the warp leaves a border of zeros
        return -10 * np.log10(((a[8:-8, 8:-8] - c[8:-8, 8:-8]) ** 2).mean())
    for b in range(B):
        for t in range(1, T):
            assert psnr_in(reg[b, t].numpy(), burst[b, 0]) > psnr_in(burst[b, t], burst[b, 0]) + 6.0


def test_ecc_degenerate_pair_reports_failure(cuda):
    """A constant frame has no gradient: J^T J is singular, cv2 throws; the kernel flags the pair (iters < 0, rho = -1) and
    leaves the other pairs alone."""
    from fbanet_b200 import ops
    x = torch.rand(1, 3, 1, 40, 48, generator=torch.Generator().manual_seed(0))
    x[0, 1] = 0.5
    x[0, 2] = x[0, 0]
    M, rho, iters = ops.ecc_homography_burst(x.to(cuda))
    assert iters[0, 1].item() < 0 and rho[0, 1].item() == -1.0
    assert iters[0, 2].item() > 0 and rho[0, 2].item() > 0.999
    assert (M[0, 2].cpu() - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-3


def test_tile_kernels_match_vectors_from_the_reference_code(cuda):
    """8f-1: the GPU tile gather / stitch (plain and row-band forms) against vectors produced by running the reference's own
    utils/dataset_utils.py (tests/golden/make_golden_reference.py) -- bit exact."""
    from fbanet_b200 import ops
    from fbanet_b200.dist import band_rows
    from tests_golden_helpers import tiling_reference
    d, sr = tiling_reference()
    ps, ov, sc = int(d["psize"]), int(d["overlap"]), int(d["scale"])
    burst = torch.from_numpy(d["burst"])[0].to(cuda)
    T, Cc, H, W = burst.shape
    assert np.array_equal(ops.tile_divide(burst, ps, ov).cpu().numpy(), d["tiles"])
    out = torch.zeros(Cc, sc * H, sc * W, device=cuda)
    ops.tile_merge(torch.from_numpy(sr).to(cuda), out, H, W, ps, ov, sc)
    assert np.array_equal(out.cpu().numpy(), d["merged"][0])
    row0 = band_rows(H, 3)
    bands = [burst[:, :, row0[k]:row0[k + 1]].contiguous() for k in range(3)]
    n = d["tiles"].shape[0]
    assert np.array_equal(ops.tile_divide_banded([b.data_ptr() for b in bands], row0, T, Cc, H, W, ps, ov, 0, n, cuda).cpu().numpy(), d["tiles"])
    ob = [torch.zeros(Cc, sc * (row0[k + 1] - row0[k]), sc * W, device=cuda) for k in range(3)]
    ops.tile_merge_banded(torch.from_numpy(sr).to(cuda), [b.data_ptr() for b in ob], row0, H, W, ps, ov, sc, 0, n)
    assert np.array_equal(torch.cat(ob, 1).cpu().numpy(), d["merged"][0])


def test_ecc_large_frame_takes_the_l2_path(cuda):
    """Frames whose two planes do not fit in shared memory (here 232 x 248: 460 KB) run the ECC iterations with their taps through
    L2 instead; same result contract against the float64 restatement."""
    import cv2
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import bgr2gray, ecc_homography, homography_coord_diff
    H, W = 232, 248
    rng = np.random.default_rng(21)
    base = cv2.GaussianBlur(rng.random((H, W, 3)).astype(np.float32), (0, 0), 3.0)
    base = (base - base.min()) / (base.max() - base.min())
    M = np.eye(3, dtype=np.float32)
    M[0, 2], M[1, 2], M[0, 1], M[2, 0] = 2.4, -1.6, 0.006, 8e-6
    frame = cv2.warpPerspective(base, M, (W, H), flags=cv2.INTER_LINEAR) + rng.normal(0, 0.01, (H, W, 3)).astype(np.float32)
    burst = torch.from_numpy(np.stack([base, frame])[None]).contiguous()          # [1,2,H,W,3]
    Mg, rho, iters = ops.ecc_homography_burst(burst.to(cuda), layout="BTHWC")
    r_ref, M_ref = ecc_homography(bgr2gray(base), bgr2gray(frame))
    assert iters[0, 1].item() > 0 and abs(rho[0, 1].item() - r_ref) < 2e-4
    assert homography_coord_diff(Mg[0, 1].cpu().numpy(), M_ref, H, W) < 5e-3


@pytest.mark.parametrize("shape,gw", [((2, 3, 14, 18), 3.0), ((1, 3, 64, 80), 3.0), ((3, 4, 33, 17), 0.0), ((1, 1, 5, 300), 1.5)])
def test_training_loss_matches_oracle_autograd(cuda, shape, gw):
    """8f-3: CharbonnierLoss + gw * GWLoss (losses.py:39-80, train.py.bak:168) -- value and dL/d(restored) from the one-pass kernel
    against torch autograd through the restated losses, on images that leave [0,1] (clamp) and share an identical patch (sign(0))."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import training_loss
    g = torch.Generator().manual_seed(9)
    x = torch.rand(shape, generator=g) * 1.3 - 0.15
    y = torch.rand(shape, generator=g) * 1.3 - 0.15
    y[0, 0, 1:4, 2:5] = x[0, 0, 1:4, 2:5]
    xr = x.clone().requires_grad_(True)
    ref = training_loss(xr, y, gw_weight=gw)
    ref.backward()
    loss, grad = ops.training_loss(x.to(cuda), y.to(cuda), gw_weight=gw)
    assert abs(loss[0].item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert (grad.cpu() - xr.grad).abs().max().item() < 1e-5 * xr.grad.abs().max().item() + 1e-9
    loss2, none = ops.training_loss(x.to(cuda), y.to(cuda), gw_weight=gw, need_grad=False)
    assert none is None and torch.equal(loss2, loss)                              # deterministic reduction


def test_training_loss_matches_vectors_from_the_reference_code(cuda):
    import os
    from fbanet_b200 import ops
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_reference.npz"))
    loss, grad = ops.training_loss(torch.from_numpy(d["x"]).to(cuda), torch.from_numpy(d["y"]).to(cuda))
    assert abs(loss[0].item() - float(d["total"])) < 1e-5 * float(d["total"]) and abs(loss[1].item() - float(d["charbonnier"])) < 1e-6
    assert abs(loss[2].item() - float(d["gw"])) < 1e-5 * float(d["gw"])
    assert np.abs(grad.cpu().numpy() - d["grad"]).max() < 1e-5 * np.abs(d["grad"]).max()
    # train.py.bak:167-168 as written: clamp(restored, 0, 1) before both criteria (what train_step computes)
    loss, grad = ops.training_loss(torch.from_numpy(d["x"]).to(cuda), torch.from_numpy(d["y"]).to(cuda), clamp_restored=True)
    assert abs(loss[0].item() - float(d["total_clamped"])) < 1e-5 * float(d["total_clamped"])
    assert np.abs(grad.cpu().numpy() - d["grad_clamped"]).max() < 1e-5 * np.abs(d["grad_clamped"]).max()


@pytest.mark.parametrize("decoupled,wd", [(True, 0.02), (False, 0.02), (True, 0.0)])
def test_adam_step_matches_torch_optim(cuda, decoupled, wd):
    """8f-3: the flat-buffer optimizer step against torch.optim.AdamW / Adam (the reference's optimizer, train.py.bak:72-78) over five
    steps with fresh gradients; CPU torch is the reference here (it is the library the reference trainer calls)."""
    from fbanet_b200 import ops
    g = torch.Generator().manual_seed(4)
    n = 100_003
    w0 = torch.randn(n, generator=g) * 0.1
    ref = torch.nn.Parameter(w0.clone())
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref], lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    w, m, v = w0.clone().to(cuda), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    for step in range(1, 6):
        gr = torch.randn(n, generator=g) * (0.5 if step % 2 else 2.0)
        ref.grad = gr.clone()
        opt.step()
        ops.adam_step(w, (gr * 4.0).to(cuda), m, v, step, lr=2e-4, weight_decay=wd, decoupled=decoupled, grad_scale=0.25)   # as after a 4-rank sum
        assert (w.cpu() - ref.detach()).abs().max().item() < 2e-7, step
    st = opt.state[ref]
    assert (m.cpu() - st["exp_avg"]).abs().max().item() < 1e-6 and (v.cpu() - st["exp_avg_sq"]).abs().max().item() < 1e-6


def test_flat_params_adam_step_matches_torch_on_a_module(cuda):
    """8f-3: FlatParams (parameters re-pointed into one buffer) + the fused AdamW step against torch.optim.AdamW on a twin module,
    three steps with autograd gradients accumulated straight into the flat gradient buffer."""
    import copy
    from fbanet_b200.train import FlatParams
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.PReLU(), torch.nn.Conv2d(8, 3, 3, padding=1)).to(cuda)
    twin = copy.deepcopy(net)
    opt = torch.optim.AdamW(twin.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    flat = FlatParams(net.parameters())
    for step in range(3):
        x = torch.rand(2, 3, 16, 16, device=cuda, generator=torch.Generator(device=cuda).manual_seed(step))
        flat.zero_grad()
        opt.zero_grad()
        net(x).square().mean().backward()
        twin(x).square().mean().backward()
        flat.adam_step(1e-3, weight_decay=0.02)
        opt.step()
        for a, b in zip(net.parameters(), twin.parameters()):
            assert (a - b).abs().max().item() < 1e-6, step
    assert set(net.state_dict()) == set(twin.state_dict())


def test_torch_library_custom_ops(cuda):
    """the C-ABI launchers as ``torch.library`` custom ops -- ``torch.ops.fbanet.*`` against torch references on
    bf16-rounded operands (conv3x3 + PReLU-free ReLU + residual, linear + GELU, LayerNorm, window attention, the LeFF MLP in both its
    one-kernel (C = 64) and two-kernel (C = 256) forms, warp), and ``fbanet::forward`` == the model's own forward."""
    import fbanet_b200.torch_ops as T
    from fbanet_b200 import BaseModel, ops, _lib as L
    from oracle.fbanet_oracle import build_oracle
    dt = torch.bfloat16
    # conv3x3 + ReLU + residual
    x, w, b, r = _r(dt, 2, 64, 24, 20, seed=1), _r(dt, 64, 64, 3, 3, seed=2, scale=0.05), _r(torch.float32, 64, seed=3), _r(dt, 2, 64, 24, 20, seed=4)
    ref = F.relu(F.conv2d(x, w, b, padding=1)) + r
    got = torch.ops.fbanet.conv3x3(_nhwc(x, dt, cuda), _pack(w, dt, cuda), b.to(cuda), _nhwc(r, dt, cuda), L.ACT_RELU)
    _close(got.permute(0, 3, 1, 2), ref, dt)
    # linear + GELU
    xl, wl, bl = _r(dt, 1, 15, 20, 128, seed=5), _r(dt, 512, 128, seed=6, scale=0.08), _r(torch.float32, 512, seed=7)
    got = torch.ops.fbanet.linear(xl.to(cuda, dt), wl.to(cuda, dt), bl.to(cuda), None, L.ACT_GELU_TANH)
    _close(got, F.gelu(F.linear(xl, wl, bl), approximate="tanh"), dt)
    # LayerNorm
    g, be = _r(torch.float32, 128, seed=8) + 1.0, _r(torch.float32, 128, seed=9)
    _close(torch.ops.fbanet.layernorm(xl.to(cuda, dt), g.to(cuda), be.to(cuda), 1e-5), F.layer_norm(xl, (128,), g, be, 1e-5), dt, scale=2.0)
    # window attention == the ctypes wrapper (itself checked against the oracle above)
    qkv, table = _r(dt, 2, 20, 20, 192, seed=10, scale=1.5).to(cuda, dt), _r(torch.float32, 361, 4, seed=11, scale=0.5).to(cuda)
    a = torch.ops.fbanet.window_attention(qkv, table, 4, 10, 5, 0.25)
    assert torch.equal(a.view(-1, 64), ops.window_attention(qkv.view(-1, 192), table, 2, 20, 20, 4, 10, 5, 0.25))
    # LeFF MLP, both forms
    for C in (64, 256):
        Hd = 4 * C
        xm, res = _r(dt, 1, C, 16, 24, seed=12), _r(dt, 1, C, 16, 24, seed=13)
        w1, b1 = _r(dt, Hd, C, seed=14, scale=1 / math.sqrt(C)), _r(torch.float32, Hd, seed=15, scale=0.2)
        dw, db = _r(torch.float32, Hd, 1, 3, 3, seed=16, scale=0.3), _r(torch.float32, Hd, seed=17, scale=0.1)
        w2, b2 = _r(dt, C, Hd, seed=18, scale=1 / math.sqrt(Hd)), _r(torch.float32, C, seed=19)
        gelu = lambda v: F.gelu(v, approximate="tanh")
        h1 = gelu(F.linear(xm.permute(0, 2, 3, 1), w1, b1)).permute(0, 3, 1, 2).to(dt).float()
        mid = gelu(F.conv2d(h1, dw, db, padding=1, groups=Hd)).to(dt).float()
        ref = F.linear(mid.permute(0, 2, 3, 1), w2, b2) + res.permute(0, 2, 3, 1)
        got = torch.ops.fbanet.leff_mlp(_nhwc(xm, dt, cuda), w1.to(cuda), b1.to(cuda), dw.to(cuda), db.to(cuda), w2.to(cuda), b2.to(cuda),
                                        _nhwc(res, dt, cuda), L.ACT_GELU_TANH)
        _close(got, ref, dt, scale=2.0)
    # warp == the ctypes wrapper; forward == the module call
    burst = torch.rand(1, 4, 3, 40, 40, device=cuda)
    M = torch.eye(3, dtype=torch.float64).repeat(1, 4, 1, 1)
    M[:, 1:, 0, 2] = 1.5
    assert torch.equal(torch.ops.fbanet.warp(burst, M), ops.warp_burst(burst, M))
    cfg = dict(num_frames=4, img_size=40, in_channels=3, embed_dim=32, window_length=10)
    m = BaseModel(**cfg, token_projection="linear", token_mlp="leff", dtype="fp32")
    m.load_state_dict(build_oracle(0, **cfg).state_dict())
    m = m.to(cuda).eval()
    h = T.register_model(m)
    assert torch.equal(torch.ops.fbanet.forward(burst, h), m(burst))


@pytest.mark.parametrize("cin,hw,bursts,T", [(3, (40, 40), 2, 4), (4, (37, 21), 1, 5), (3, (160, 160), 2, 14), (4, (80, 80), 3, 14)])
def test_head_conv_with_fused_homography_warp(cuda, cin, hw, bursts, T):
    """K1 fused into K0 (homography_alignment.py:46-55 -> models/fba_net.py:255): the head conv sampling the unregistered burst through
    the homographies equals warp kernel + head conv BIT FOR BIT (same fp64 coordinates, same tap order), base frames copied,
    strong perspective / rotation so taps leave the image on every side, ragged tiles; and the warp itself is pinned to the
    float64 oracle (1e-5 px coordinates) by test_warp_matches_float64_oracle."""
    from fbanet_b200 import ops
    H, W = hw
    g = torch.Generator().manual_seed(5)
    x = torch.rand(bursts, T, cin, H, W, generator=g).to(cuda)
    M = torch.eye(3, dtype=torch.float64).repeat(bursts, T, 1, 1)
    M[:, 1:, :2, 2] = torch.rand(bursts, T - 1, 2, generator=g, dtype=torch.float64) * 12 - 6
    M[:, 1:, :2, :2] += torch.rand(bursts, T - 1, 2, 2, generator=g, dtype=torch.float64) * 0.2 - 0.1
    M[:, 1:, 2, :2] = torch.rand(bursts, T - 1, 2, generator=g, dtype=torch.float64) * 2e-4 - 1e-4
    w, b = _r(torch.float32, 64, cin, 3, 3, seed=2, scale=0.2), _r(torch.float32, 64, seed=3)
    wkc = w.permute(2, 3, 1, 0).reshape(9 * cin, 64).contiguous().to(cuda)
    ref = ops.head_conv(ops.warp_burst(x, M).view(bursts * T, cin, H, W), wkc, b.to(cuda), torch.bfloat16)
    launches = ops.LAUNCHES
    got = ops.head_conv(x.view(bursts * T, cin, H, W), wkc, b.to(cuda), torch.bfloat16, M=M.view(-1, 3, 3), frames_per_burst=T)
    assert ops.LAUNCHES == launches + 1          # one kernel: no warp launch, no fallback
    # the planar warp kernel shares one fp64 division between four pixels (coordinates equal to ~1e-13 px), so a bilinear weight may
    # round differently once in a few million samples: allow a handful of features to move by an ulp
    diff = (got.float() - ref.float()).abs()
    assert (diff > 0).float().mean().item() < 1e-5 and diff.max().item() < 0.02, ((diff > 0).sum().item(), diff.max().item())


def test_model_forward_with_homographies_fuses_the_warp(cuda):
    """``model(burst, homographies=M)`` (cfg3's front end) == ``model(warp_burst(burst, M))``: bit for bit in the default mode (warp
    kernel, then the forward), and with the warp fused into the head conv (``fuse_warp``, tensor-core path: one launch fewer, the same
    image up to a weight rounding in a few samples per million); the fp32 path warps first either way."""
    from fbanet_b200 import BaseModel, ops
    from oracle.fbanet_oracle import build_oracle
    cfg = dict(num_frames=5, img_size=40, in_channels=4, embed_dim=64, window_length=10)
    sd = build_oracle(0, **cfg).state_dict()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 5, 4, 40, 40, generator=g).to(cuda)
    M = torch.eye(3, dtype=torch.float64).repeat(2, 5, 1, 1)
    M[:, 1:, :2, 2] = torch.rand(2, 4, 2, generator=g, dtype=torch.float64) * 6 - 3
    M[:, 1:, 2, :2] = torch.rand(2, 4, 2, generator=g, dtype=torch.float64) * 2e-5 - 1e-5
    for dtype in ("bf16", "fp32"):
        m = BaseModel(**cfg, token_projection="linear", token_mlp="leff", dtype=dtype)
        m.load_state_dict(sd)
        m = m.to(cuda).eval()
        ref = m(ops.warp_burst(x, M))
        n0 = ops.LAUNCHES
        m(ops.warp_burst(x, M))
        n_two = ops.LAUNCHES - n0
        for fuse in (False, True):
            m.fuse_warp = fuse
            n0 = ops.LAUNCHES
            got = m(x, homographies=M)
            n = ops.LAUNCHES - n0
            assert (got - ref).abs().max().item() < 1e-3, (dtype, fuse)
            if not fuse:
                assert torch.equal(got, ref), dtype
            assert n == (n_two - 1 if (dtype == "bf16" and fuse) else n_two), (dtype, fuse, n, n_two)
