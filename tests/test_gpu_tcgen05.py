"""tcgen05/TMA implicit-GEMM kernel (bf16) against torch CPU fp32 references on bf16-rounded operands.
``impl=IMPL_TCGEN05`` is forced so a silent SIMT fallback cannot make these pass."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g) * 2 - 1) * scale).to(BF).float()


def _nhwc(x, dev):
    return x.permute(0, 2, 3, 1).contiguous().to(dev, BF)


def _pack(w, dev):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(dev, BF)


def _check(got, ref, tol=2e-2):
    got, ref = got.float().cpu(), ref.float()
    err = (got - ref).abs().max().item()
    assert torch.allclose(got, ref, atol=tol, rtol=tol), f"max abs err {err}"


@pytest.mark.parametrize("cin,cout,h,w,act,res", [
    (64, 64, 160, 160, 1, False), (64, 64, 24, 40, 0, True), (128, 128, 80, 80, 0, True), (256, 256, 40, 40, 1, False),
    (128, 64, 20, 20, 2, False), (64, 256, 10, 10, 3, False), (256, 128, 33, 17, 0, True), (64, 64, 5, 7, 0, False)])
def test_tc_conv3x3(cuda, cin, cout, h, w, act, res):
    from fbanet_b200 import ops, _lib as L
    n = 3 if h <= 40 else 1
    x, wt, b = _r(n, cin, h, w, seed=1), _r(cout, cin, 3, 3, seed=2, scale=1 / math.sqrt(cin * 9)), _r(cout, seed=3, scale=0.1)
    r = _r(n, cout, h, w, seed=4)
    alpha = torch.tensor([0.25])
    ref = F.conv2d(x, wt, b, padding=1)
    ref = {0: lambda v: v, 1: F.relu, 2: lambda v: F.prelu(v, alpha), 3: lambda v: F.gelu(v, approximate="tanh")}[act](ref)
    if res:
        ref = ref + r
    out = torch.empty(n, h, w, cout, device=cuda, dtype=BF)
    ops.conv_gemm([_nhwc(x, cuda)], _pack(wt, cuda), out, kh=3, kw=3, pad=1, bias=b.to(cuda), act=act, alpha=alpha.to(cuda),
                  residual=_nhwc(r, cuda) if res else None, impl=L.IMPL_TCGEN05)
    _check(out.permute(0, 3, 1, 2), ref)


def test_tc_multisource_views_and_slice_output(cuda):
    from fbanet_b200 import ops, _lib as L
    a, b_, c = _r(2, 256, 20, 20, seed=1), _r(2, 128, 20, 20, seed=2), _r(2, 128, 20, 20, seed=3)
    w, bias = _r(256, 512, 3, 3, seed=4, scale=0.02), _r(256, seed=5)
    ref = F.prelu(F.conv2d(torch.cat([a, b_, c], 1), w, bias, padding=1), torch.tensor([0.25]))
    wide = torch.zeros(2, 20, 20, 192, device=cuda, dtype=BF)
    wide[..., 64:] = _nhwc(c, cuda)
    outbuf = torch.zeros(2, 20, 20, 512, device=cuda, dtype=BF)
    ops.conv_gemm([_nhwc(a, cuda), _nhwc(b_, cuda), wide[..., 64:]], _pack(w, cuda), outbuf[..., 256:], kh=3, kw=3, pad=1, bias=bias.to(cuda),
                  act=L.ACT_PRELU, alpha=torch.tensor([0.25], device=cuda), impl=L.IMPL_TCGEN05)
    _check(outbuf[..., 256:].permute(0, 3, 1, 2), ref)
    assert outbuf[..., :256].abs().max().item() == 0


@pytest.mark.parametrize("cin,cout", [(64, 192), (128, 384), (256, 768), (128, 512), (256, 1024), (512, 128), (1024, 256), (896, 64)])
def test_tc_linear(cuda, cin, cout):
    """token GEMMs: qkv (N=3C, BN 192/256 tiling), LeFF fc1/fc2, FAF 1x1 (K=896)."""
    from fbanet_b200 import ops, _lib as L
    T = 1700
    x, w, b = _r(1, T, cin, seed=1), _r(cout, cin, seed=2, scale=1 / math.sqrt(cin)), _r(cout, seed=3)
    r = _r(1, T, cout, seed=4)
    ref = F.linear(x, w, b) + r
    out = torch.empty(1, 34, 50, cout, device=cuda, dtype=BF)
    ops.conv_gemm([x.view(1, 34, 50, cin).to(cuda, BF)], w.to(cuda, BF), out, bias=b.to(cuda), residual=r.view(1, 34, 50, cout).to(cuda, BF),
                  impl=L.IMPL_TCGEN05)
    _check(out.view(1, T, cout), ref, tol=3e-2)


@pytest.mark.parametrize("cin,co,hw", [(256, 128, 40), (256, 64, 80), (256, 128, 10)])
def test_tc_conv_transpose(cuda, cin, co, hw):
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(2, cin, hw, hw, seed=1), _r(cin, co, 2, 2, seed=2, scale=0.05), _r(co, seed=3)
    ref = F.conv_transpose2d(x, w, b, stride=2)
    wp = w.permute(2, 3, 1, 0).reshape(4 * co, cin).contiguous().to(cuda, BF)
    cat = torch.zeros(2, 2 * hw, 2 * hw, 2 * co, device=cuda, dtype=BF)
    ops.conv_gemm([_nhwc(x, cuda)], wp, cat[..., :co], bias=b.repeat(4).to(cuda), store_mode=L.STORE_CONVT2, impl=L.IMPL_TCGEN05)
    _check(cat[..., :co].permute(0, 3, 1, 2), ref)
    assert cat[..., co:].abs().max().item() == 0


@pytest.mark.parametrize("cin,cout,hw", [(64, 128, 160), (128, 256, 80), (64, 128, 20)])
def test_tc_downsample_space_to_depth(cuda, cin, cout, hw):
    from fbanet_b200 import ops, _lib as L
    n = 2
    x, w, b = _r(n, cin, hw, hw, seed=1), _r(cout, cin, 4, 4, seed=2, scale=1 / math.sqrt(16 * cin)), _r(cout, seed=3)
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    wide = torch.zeros(n, hw, hw, 2 * cin, device=cuda, dtype=BF)
    wide[..., cin:] = _nhwc(x, cuda)  # strided source view, as conv0 inside the concat buffer
    s2d = ops.space_to_depth(wide[..., cin:])
    xs = x.permute(0, 2, 3, 1).reshape(n, hw // 2, 2, hw // 2, 2, cin).permute(0, 1, 3, 2, 4, 5).reshape(n, hw // 2, hw // 2, 4 * cin)
    assert torch.equal(s2d.float().cpu(), xs)
    out = torch.empty(n, hw // 2, hw // 2, cout, device=cuda, dtype=BF)
    ops.conv_gemm([s2d], _pack(w, cuda), out, kh=4, kw=4, stride=2, pad=1, bias=b.to(cuda), src_s2d=True, impl=L.IMPL_TCGEN05)
    _check(out.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("cin,cout,h,w,n", [(64, 128, 32, 32, 2), (128, 256, 16, 16, 2), (64, 128, 160, 160, 2), (64, 128, 20, 44, 3), (128, 256, 6, 10, 1), (64, 64, 2, 2, 2)])
def test_tc_downsample_element_strided_boxes(cuda, cin, cout, h, w, n):
    """The 4x4 stride-2 downsample conv straight from its full-resolution input: every tap is a TMA box that steps two pixels
    (elementStrides {1,2,2,1}; layers/downsample.py) -- no space-to-depth copy.  Strided source view inside a concat buffer, ragged
    tiles, and the same bits as the space-to-depth form (same MMAs on the same operands in the same order)."""
    from fbanet_b200 import ops, _lib as L
    x, wt, b = _r(n, cin, h, w, seed=1), _r(cout, cin, 4, 4, seed=2, scale=1 / math.sqrt(16 * cin)), _r(cout, seed=3)
    ref = F.conv2d(x, wt, b, stride=2, padding=1)
    wide = torch.zeros(n, h, w, 2 * cin, device=cuda, dtype=BF)
    wide[..., cin:] = _nhwc(x, cuda)
    src = wide[..., cin:]
    out = torch.empty(n, h // 2, w // 2, cout, device=cuda, dtype=BF)
    ops.conv_gemm([src], _pack(wt, cuda), out, kh=4, kw=4, stride=2, pad=1, bias=b.to(cuda), impl=L.IMPL_TCGEN05)
    _check(out.permute(0, 3, 1, 2), ref)
    out2 = torch.empty_like(out)
    ops.conv_gemm([ops.space_to_depth(src)], _pack(wt, cuda), out2, kh=4, kw=4, stride=2, pad=1, bias=b.to(cuda), src_s2d=True, impl=L.IMPL_TCGEN05)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("cin,ct,h,w,n", [(64, 2, 160, 160, 2), (64, 1, 13, 29, 3), (128, 3, 40, 22, 2), (64, 4, 6, 14, 1), (192, 2, 7, 15, 2)])
def test_tc_tapsum_conv3x3(cuda, cin, ct, h, w, n):
    """3x3 convs with <= 4 fp32 outputs run tap-stacked (nine taps along N, shifted sum in the epilogue): must equal the plain
    convolution, including images that are not multiples of the 14 x 6 interior, several K chunks, and the FBANET_TC_TAPSUM=0
    fallback (same result from the halo-mode implicit GEMM)."""
    import os
    from fbanet_b200 import ops, _lib as L
    x, wt, b = _r(n, cin, h, w, seed=1), _r(ct, cin, 3, 3, seed=2, scale=1 / math.sqrt(cin * 9)), _r(ct, seed=3, scale=0.1)
    ref = F.conv2d(x, wt, b, padding=1)
    wp = F.pad(_pack(wt, cuda), (0, 0, 0, 16 - ct)).contiguous()        # N padded to the tensor-core minimum, as the model packs it
    bp = F.pad(b, (0, 16 - ct)).to(cuda)
    outs = []
    for mode in ("1", "0"):
        os.environ["FBANET_TC_TAPSUM"] = mode
        out = torch.full((n, h, w, ct), -7.0, device=cuda, dtype=torch.float32)   # odd widths (1, 3) take the halo path in both modes
        ops.conv_gemm([_nhwc(x, cuda)], wp, out, kh=3, kw=3, pad=1, bias=bp, store_mode=L.STORE_NHWC_F32, cout_store=ct, impl=L.IMPL_TCGEN05)
        _check(out.permute(0, 3, 1, 2), ref, tol=1e-2)
        outs.append(out.cpu())
    os.environ.pop("FBANET_TC_TAPSUM")
    assert (outs[0] - outs[1]).abs().max() < 1e-4                        # same products, different summation order


def test_tc_final_conv_nchw_base(cuda):
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(2, 64, 64, 96, seed=1), _r(3, 64, 3, 3, seed=2, scale=0.03), _r(3, seed=3)
    burst = torch.rand(2, 5, 3, 16, 24, generator=torch.Generator().manual_seed(4))
    ref = F.conv2d(x, w, b, padding=1) + F.interpolate(burst[:, 0], scale_factor=4, mode="bilinear", align_corners=False)
    wp = F.pad(_pack(w, cuda), (0, 0, 0, 13))
    bp = F.pad(b, (0, 13)).to(cuda)
    out = torch.empty(2, 3, 64, 96, device=cuda, dtype=torch.float32)
    ops.conv_gemm([_nhwc(x, cuda)], wp.contiguous(), out, kh=3, kw=3, pad=1, bias=bp, store_mode=L.STORE_NCHW_BASE, base=burst.to(cuda)[:, 0],
                  cout_store=3, impl=L.IMPL_TCGEN05)
    _check(out, ref, tol=1e-2)


def test_tc_head_im2col(cuda):
    from fbanet_b200 import ops, _lib as L
    x = torch.rand(6, 3, 40, 40, generator=torch.Generator().manual_seed(1))
    w, b = _r(64, 3, 3, 3, seed=2, scale=0.2), _r(64, seed=3)
    ref = F.conv2d(x.to(BF).float(), w, b, padding=1)
    xn = ops.to_nhwc(x.to(cuda), 64, BF, im2col3x3=True)
    wp = F.pad(w.permute(0, 2, 3, 1).reshape(64, 27), (0, 37)).to(cuda, BF).contiguous()
    out = torch.empty(6, 40, 40, 64, device=cuda, dtype=BF)
    ops.conv_gemm([xn], wp, out, bias=b.to(cuda), impl=L.IMPL_TCGEN05)
    _check(out.permute(0, 3, 1, 2), ref)


def test_tc_gated_features_and_fuse(cuda):
    """K2 on the tensor-core path: gate kernel emits pixel-major gated features, 1x1 fusion GEMM K=F*E."""
    from fbanet_b200 import ops, _lib as L
    B, Fr, E, S = 2, 14, 64, 20
    feat = _r(B, Fr, S, S, E, seed=1)
    w1 = _r(E, E, 3, 3, seed=4, scale=0.04)
    wsum = w1.double().sum(0).permute(1, 2, 0).reshape(9, E).float().contiguous()
    fd = feat.to(cuda, BF)
    gate, gated = ops.faf_gate(fd, wsum.to(cuda), want_gate=True, want_gated=True)
    g = gate.cpu()
    exp = torch.cat([feat[:, :1], feat[:, 1:] * g[..., None]], 1).permute(0, 2, 3, 1, 4).reshape(B, S, S, Fr * E)
    _check(gated, exp, tol=1e-2)
    w, b = _r(E, Fr * E, seed=5, scale=0.03), _r(E, seed=6)
    ref = F.prelu(F.linear(gated.float().cpu(), w, b), torch.tensor([0.1]))
    out = torch.empty(B, S, S, E, device=cuda, dtype=BF)
    ops.conv_gemm([gated], w.to(cuda, BF), out, bias=b.to(cuda), act=L.ACT_PRELU, alpha=torch.tensor([0.1], device=cuda), impl=L.IMPL_TCGEN05)
    _check(out, ref)


def test_tc_persistent_many_tiles(cuda):
    """more tiles than SMs and than pipeline stages: exercises the persistent loop, barrier phase wrap and
    TMEM double buffering (batch 24 x 80x80 -> 1200 M-tiles x 2 N-tiles)."""
    from fbanet_b200 import ops, _lib as L
    x, w, b = _r(24, 128, 80, 80, seed=1), _r(512, 128, seed=2, scale=0.08), _r(512, seed=3)
    ref = F.gelu(F.linear(x.permute(0, 2, 3, 1), w, b), approximate="tanh")
    out = torch.empty(24, 80, 80, 512, device=cuda, dtype=BF)
    ops.conv_gemm([_nhwc(x, cuda)], w.to(cuda, BF), out, bias=b.to(cuda), act=L.ACT_GELU_TANH, impl=L.IMPL_TCGEN05)
    _check(out, ref)


def test_unsupported_shapes_are_refused(cuda):
    from fbanet_b200 import ops, _lib as L
    x = torch.zeros(1, 8, 8, 32, device=cuda, dtype=BF)
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.conv_gemm([x], torch.zeros(64, 32, device=cuda, dtype=BF), torch.empty(1, 8, 8, 64, device=cuda, dtype=BF), impl=L.IMPL_TCGEN05)


@pytest.mark.parametrize("C,Hd,n,h,w", [(64, 256, 2, 32, 24), (128, 512, 2, 16, 40), (256, 1024, 1, 40, 40), (128, 512, 3, 20, 20), (64, 256, 70, 16, 16)])
def test_fused_leff_dwconv_fc2(cuda, C, Hd, n, h, w):
    """LeFF tail fused on the tensor cores: Linear2(GELU(dwconv3x3(h1) + b)) + b2 + residual
    (locally_enhanced_feed_forward.py:39-57); the last case has more tiles than SMs (persistent loop, phase wrap)."""
    from fbanet_b200 import ops, _lib as L
    h1 = _r(n, Hd, h, w, seed=1)
    dw, db = _r(Hd, 1, 3, 3, seed=2, scale=0.3), _r(Hd, seed=3, scale=0.1)
    w2, b2 = _r(C, Hd, seed=4, scale=1 / math.sqrt(Hd)), _r(C, seed=5)
    res = _r(n, C, h, w, seed=6)
    mid = F.gelu(F.conv2d(h1, dw, db, padding=1, groups=Hd), approximate="tanh").to(BF).float()
    ref = F.linear(mid.permute(0, 2, 3, 1), w2, b2) + res.permute(0, 2, 3, 1)
    buf = torch.zeros(n, h, w, 2 * C, device=cuda, dtype=BF)  # output into a channel slice, as the concat buffers do
    out = ops.leff_fc2(_nhwc(h1, cuda), dw.reshape(Hd, 9).t().contiguous().to(cuda), db.to(cuda), w2.to(cuda, BF), b2.to(cuda), buf[..., C:],
                       _nhwc(res, cuda), L.ACT_GELU_TANH)
    assert out is not None
    _check(buf[..., C:], ref, tol=3e-2)
    assert buf[..., :C].abs().max().item() == 0


@pytest.mark.parametrize("C,Hd,n,h,w", [(256, 1024, 2, 16, 24), (256, 1024, 1, 37, 21), (128, 512, 2, 20, 20), (256, 1024, 3, 40, 40)])
def test_fc1_fp16_store_and_fused_leff_tail_on_half2(cuda, C, Hd, n, h, w):
    """The dim-256 LeFF in fp16: the fc1 GEMM's staged epilogue stores its GELU output as IEEE fp16 (`conv_gemm` with an fp16 `out`) and
    `leff_fc2` with fp16 h1 / fp16 fc2 weights runs the depthwise conv + GELU on packed half2 with an fp16 A tile.  Against torch fp32
    with the hidden map rounded to fp16 where the kernels round it."""
    from fbanet_b200 import ops, _lib as L
    x = _r(n, C, h, w, seed=1)
    w1, b1 = _r(Hd, C, seed=2, scale=1 / math.sqrt(C)), _r(Hd, seed=3, scale=0.2)
    dw, db = _r(Hd, 1, 3, 3, seed=4, scale=0.3), _r(Hd, seed=5, scale=0.1)
    w2 = ((torch.rand(C, Hd, generator=torch.Generator().manual_seed(6)) * 2 - 1) / math.sqrt(Hd)).half().float()
    b2, res = _r(C, seed=7), _r(n, C, h, w, seed=8)
    gelu = lambda v: F.gelu(v, approximate="tanh")
    h1_ref = gelu(F.linear(x.permute(0, 2, 3, 1), w1, b1)).half()
    h1 = torch.empty(n, h, w, Hd, device=cuda, dtype=torch.float16)
    ops.conv_gemm([_nhwc(x, cuda)], w1.to(cuda, BF), h1, bias=b1.to(cuda), act=L.ACT_GELU_TANH, impl=L.IMPL_TCGEN05)
    e = (h1.float().cpu() - h1_ref.float()).abs()
    assert (e <= 2e-3 + 4e-3 * h1_ref.float().abs()).all(), e.max().item()      # fp16 rounding + tanh.approx
    mid = gelu(F.conv2d(h1.float().cpu().permute(0, 3, 1, 2), dw, db, padding=1, groups=Hd)).half().float()
    ref = F.linear(mid.permute(0, 2, 3, 1), w2, b2) + res.permute(0, 2, 3, 1)
    out = torch.empty(n, h, w, C, device=cuda, dtype=BF)
    r = ops.leff_fc2(h1, dw.reshape(Hd, 9).t().contiguous().to(cuda), db.to(cuda), w2.to(cuda, torch.float16), b2.to(cuda), out, _nhwc(res, cuda),
                     L.ACT_GELU_TANH)
    assert r is not None
    got = out.float().cpu()
    err = (got - ref).abs()
    assert torch.isfinite(got).all() and (err <= 3e-2 + 2e-2 * ref.abs()).all(), err.max().item()


@pytest.mark.parametrize("C,n,h,w,act", [(64, 2, 32, 24, 3), (128, 2, 16, 40, 3), (128, 3, 20, 20, 3), (64, 70, 16, 16, 3), (128, 1, 37, 21, 3),
                                         (64, 1, 160, 160, 3), (128, 2, 80, 80, 3), (64, 2, 24, 16, 4)])
@pytest.mark.parametrize("poly", [0, 1])
def test_fused_leff_mlp_one_kernel(cuda, monkeypatch, C, n, h, w, act, poly):
    """The whole LeFF MLP in one kernel -- Linear1 + GELU (recomputed on each tile's one-pixel halo) -> depthwise 3x3 + GELU ->
    Linear2 + residual (locally_enhanced_feed_forward.py:25-57, layers/fba_net.py:248) -- against torch fp32 on bf16-rounded
    operands: every stage shape of the model with C <= 128 (enc0 64 @160, enc1 128 @80), ragged sizes (37 x 21: partial tiles in
    both directions, out-of-image halo everywhere), more tiles than SMs (persistent loop, barrier phase wrap), views with a row
    stride (concat slices), the erf flavour, and the FMA-pipe GELU variant of the second MMA tile (FBANET_LEFF_POLY=1)."""
    from fbanet_b200 import ops, _lib as L
    if poly and act == 4:
        pytest.skip("the polynomial variant exists for tanh-GELU only")
    monkeypatch.setenv("FBANET_LEFF_POLY", str(poly))
    Hd = 4 * C
    x = _r(n, C, h, w, seed=1)
    w1, b1 = _r(Hd, C, seed=2, scale=1 / math.sqrt(C)), _r(Hd, seed=3, scale=0.2)
    dw, db = _r(Hd, 1, 3, 3, seed=4, scale=0.3), _r(Hd, seed=5, scale=0.1)
    w2, b2 = _r(C, Hd, seed=6, scale=1 / math.sqrt(Hd)), _r(C, seed=7)
    res = _r(n, C, h, w, seed=8)
    gelu = (lambda v: F.gelu(v, approximate="tanh")) if act == 3 else F.gelu
    h1 = gelu(F.linear(x.permute(0, 2, 3, 1), w1, b1)).permute(0, 3, 1, 2).to(BF).float()   # the on-chip hidden tile is bf16
    mid = gelu(F.conv2d(h1, dw, db, padding=1, groups=Hd)).to(BF).float()            # the A operand of Linear2 is bf16
    ref = F.linear(mid.permute(0, 2, 3, 1), w2, b2) + res.permute(0, 2, 3, 1)
    xin = torch.zeros(n, h, w, 2 * C, device=cuda, dtype=BF)                        # input and output as channel slices of wider buffers
    xin[..., C:] = _nhwc(x, cuda)
    buf = torch.zeros(n, h, w, 2 * C, device=cuda, dtype=BF)
    out = ops.leff_mlp(xin[..., C:], (0.5 * w1).to(cuda, BF), (0.5 * b1).to(cuda), (0.5 * dw).reshape(Hd, 9).t().contiguous().to(cuda),
                       (0.5 * db).to(cuda), w2.to(cuda, BF), b2.to(cuda), buf[..., :C], _nhwc(res, cuda), act)
    assert out is not None
    torch.cuda.synchronize()
    _check(buf[..., :C], ref, tol=3e-2)
    assert buf[..., C:].abs().max().item() == 0
    # no residual, contiguous tensors
    out2 = torch.empty(n, h, w, C, device=cuda, dtype=BF)
    assert ops.leff_mlp(_nhwc(x, cuda), (0.5 * w1).to(cuda, BF), (0.5 * b1).to(cuda), (0.5 * dw).reshape(Hd, 9).t().contiguous().to(cuda),
                        (0.5 * db).to(cuda), w2.to(cuda, BF), b2.to(cuda), out2, None, act) is not None
    _check(out2, ref - res.permute(0, 2, 3, 1), tol=3e-2)


@pytest.mark.parametrize("C,n,h,w", [(64, 2, 32, 24), (128, 2, 16, 40), (128, 3, 20, 20), (64, 70, 16, 16), (128, 1, 37, 21), (64, 1, 160, 160), (128, 2, 80, 80)])
def test_fused_leff_mlp_fp16_hidden_tile(cuda, C, n, h, w):
    """The one-kernel LeFF MLP with fp16 fc2 weights: the on-chip hidden tile and the A operand of Linear2 are fp16 and the depthwise
    3x3 + GELU run on packed half2 (HFMA2, MUFU.TANH.F16x2).  Against torch fp32 with the hidden map rounded to fp16 at the same two
    places; tolerance = the bf16 rounding of the output plus the half2 accumulation of nine taps.  Large pre-activations (|x| up to
    ~30) exercise the half2 GELU's overflow-to-inf -> tanh = 1 branch."""
    from fbanet_b200 import ops, _lib as L
    Hd = 4 * C
    x = _r(n, C, h, w, seed=1)
    w1, b1 = _r(Hd, C, seed=2, scale=1 / math.sqrt(C)), _r(Hd, seed=3, scale=0.2)
    w1[:8] *= 40.0                                                                     # a few channels with huge pre-activations
    dw, db = _r(Hd, 1, 3, 3, seed=4, scale=0.3), _r(Hd, seed=5, scale=0.1)
    w2 = ((torch.rand(C, Hd, generator=torch.Generator().manual_seed(6)) * 2 - 1) / math.sqrt(Hd)).half().float()
    b2 = _r(C, seed=7)
    res = _r(n, C, h, w, seed=8)
    gelu = lambda v: F.gelu(v, approximate="tanh")
    h1 = gelu(F.linear(x.permute(0, 2, 3, 1), w1.to(BF).float(), b1)).permute(0, 3, 1, 2).half().float()
    mid = gelu(F.conv2d(h1, dw, db, padding=1, groups=Hd)).half().float()
    ref = F.linear(mid.permute(0, 2, 3, 1), w2, b2) + res.permute(0, 2, 3, 1)
    out = torch.empty(n, h, w, C, device=cuda, dtype=BF)
    r = ops.leff_mlp(_nhwc(x, cuda), (0.5 * w1).to(cuda, BF), (0.5 * b1).to(cuda), (0.5 * dw).reshape(Hd, 9).t().contiguous().to(cuda),
                     (0.5 * db).to(cuda), w2.to(cuda, torch.float16), b2.to(cuda), out, _nhwc(res, cuda), L.ACT_GELU_TANH)
    assert r is not None
    torch.cuda.synchronize()
    got = out.float().cpu()
    err = (got - ref).abs()
    assert torch.isfinite(got).all()
    assert (err <= 3e-2 + 2e-2 * ref.abs()).all(), err.max().item()


def test_fused_leff_mlp_refuses_wide_layers(cuda):
    """C = 256 (hidden 1024) does not fit the one-kernel plan (x tile 128 KB): the op says so and the model keeps fc1 + leff_fc2."""
    from fbanet_b200 import ops, _lib as L
    C, Hd = 256, 1024
    z = lambda *s: torch.zeros(*s, device=cuda)
    assert ops.leff_mlp(z(1, 8, 8, C).to(BF), z(Hd, C).to(BF), z(Hd), z(9, Hd), z(Hd), z(C, Hd).to(BF), z(C), z(1, 8, 8, C).to(BF), None,
                        L.ACT_GELU_TANH) is None


@pytest.mark.parametrize("B,Fr,h,w", [(1, 14, 32, 24), (2, 5, 16, 40), (1, 14, 37, 21), (3, 2, 20, 20), (9, 14, 48, 32), (1, 3, 160, 160)])
def test_faf_fuse_one_pass(cuda, B, Fr, h, w):
    """K2 in one kernel -- scores, gates and the K = F*64 1x1 fusion conv + PReLU from one read of the features
    (blocks/federated_affinity_fusion.py:79-105,121-128) -- against the as-written maths in torch fp32 on bf16-rounded operands:
    gate = sigmoid(|wsum (*) f_i - wsum (*) f_0|), z = PReLU(W . cat(f_0, g_i f_i) + b).  Ragged sizes (partial tiles, out-of-image
    halo), odd frame counts (the two pixel-warp sets get unequal shares), more tiles than SMs, output into a channel slice."""
    from fbanet_b200 import ops
    E = 64
    feat = _r(B, Fr, E, h, w, seed=1)
    wsum = (_r(9, E, seed=2, scale=0.05)).float()
    hi = wsum.to(BF).float()
    wsum_used = hi + (wsum - hi).to(BF).float()                       # what the hi + lo bf16 rows carry
    wf, bf_, alpha = _r(E, Fr * E, seed=3, scale=1 / math.sqrt(Fr * E)), _r(E, seed=4, scale=0.2), torch.tensor([0.1])
    kern = wsum_used.t().reshape(1, E, 3, 3)                          # [tap, c] -> [1, c, ky, kx]
    s = F.conv2d(feat.reshape(B * Fr, E, h, w), kern, padding=1).reshape(B, Fr, h, w)
    gate = torch.sigmoid((s[:, 1:] - s[:, :1]).abs())
    gated = torch.cat([feat[:, :1], feat[:, 1:] * gate[:, :, None]], 1).reshape(B, Fr * E, h, w)
    ref = F.prelu(F.conv2d(gated, wf[:, :, None, None], bf_), alpha).permute(0, 2, 3, 1)
    fd = feat.permute(0, 1, 3, 4, 2).contiguous().to(cuda, BF)       # [B,F,H,W,E]
    buf = torch.zeros(B, h, w, 2 * E, device=cuda, dtype=BF)
    r = ops.faf_fuse(fd, ops.faf_fuse_score_weight(wsum.to(cuda)), wf.to(cuda, BF), bf_.to(cuda), alpha.to(cuda), buf[..., E:], want_gate=True)
    assert r is not None
    out, g = r
    torch.cuda.synchronize()
    assert (g.cpu() - gate).abs().max().item() < 2e-3, (g.cpu() - gate).abs().max().item()
    _check(buf[..., E:], ref, tol=3e-2)
    assert buf[..., :E].abs().max().item() == 0


@pytest.mark.parametrize("C,cout,n,h,w,act", [(64, 192, 2, 40, 40, 0), (128, 384, 1, 37, 21, 0), (256, 768, 2, 20, 20, 0), (256, 1024, 1, 40, 40, 3),
                                              (128, 384, 3, 160, 160, 0), (64, 192, 70, 16, 16, 0), (128, 512, 1, 30, 50, 3), (256, 768, 40, 40, 40, 0)])
def test_tc_linear_with_layernorm_in_shared_memory(cuda, C, cout, n, h, w, act):
    """LayerNorm applied to the A tile of a 1x1 GEMM in shared memory (norm1 -> qkv, norm2 -> fc1; layers/fba_net.py:196,246):
    the result equals the LayerNorm kernel followed by the plain GEMM BIT FOR BIT (the in-kernel LayerNorm repeats that kernel's
    arithmetic operation for operation), for every stage width (64 / 128 / 256 channels: 1, 2, 4 K chunks; resident and streamed
    weights), ragged sizes (partial tiles), more tiles than SMs (ring and barrier phase wrap), rows with a large common mode, and
    the GELU epilogue; and LayerNorm itself is pinned to torch by test_layernorm."""
    from fbanet_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(C + cout + h)
    x = ((torch.rand(n, h, w, C, generator=g) * 2 - 1) * 1.5 + torch.rand(n, h, w, 1, generator=g) * 6 - 3).to(cuda, BF)
    gam, bet = (torch.rand(C, generator=g) + 0.5).to(cuda), (torch.rand(C, generator=g) - 0.5).to(cuda)
    wt = ((torch.rand(cout, C, generator=g) * 2 - 1) / math.sqrt(C)).to(cuda, BF)
    b = (torch.rand(cout, generator=g) - 0.5).to(cuda)
    ln = ops.layernorm(x.view(-1, C), gam, bet).view(n, h, w, C)
    ref = ops.conv_gemm([ln], wt, torch.empty(n, h, w, cout, device=cuda, dtype=BF), bias=b, act=act, impl=L.IMPL_TCGEN05)
    launches = ops.LAUNCHES
    got = ops.conv_gemm([x], wt, torch.empty(n, h, w, cout, device=cuda, dtype=BF), bias=b, act=act, impl=L.IMPL_TCGEN05, ln=(gam, bet))
    assert ops.LAUNCHES == launches + 1
    assert torch.equal(got, ref), (got.float() - ref.float()).abs().max().item()
    # and against torch fp32 on the same operands (the GEMM reads bf16(LN(x)))
    t = F.layer_norm(x.float().cpu(), (C,), gam.cpu(), bet.cpu(), 1e-5).to(BF).float()
    y = F.linear(t, wt.float().cpu(), b.cpu())
    if act == 3:
        y = F.gelu(y, approximate="tanh")
    _check(got, y, tol=3e-2)


def test_model_layernorm_in_gemm_is_bit_identical(cuda):
    """The model with LayerNorm inside the qkv / fc1 GEMMs equals the model with the stand-alone LayerNorm kernel bit for bit on the
    full cfg2 shape: in the default mode (GEMMs that normalise each A tile once: the four dim-64 qkv projections with one N tile and the
    eight dim-128 ones, whose CTAs walk the three N tiles of each A tile) and with every GEMM the kernel takes
    (`ln_in_gemm = "all"`: 20 norm1 + the 8 norm2 of the dim-256 layers = 28 LayerNorm launches fewer)."""
    from fbanet_b200 import BaseModel, ops
    from oracle.fbanet_oracle import build_oracle
    cfg = dict(num_frames=14, img_size=160, in_channels=3, embed_dim=64, window_length=10)
    m = BaseModel(**cfg, token_projection="linear", token_mlp="leff", dtype="bf16")
    m.load_state_dict(build_oracle(2, **cfg).state_dict())
    m = m.to(cuda).eval()
    x = torch.rand(2, 14, 3, 160, 160, generator=torch.Generator().manual_seed(3)).to(cuda)

    def run(mode):
        m.ln_in_gemm = mode
        y = m(x)
        n0 = ops.LAUNCHES
        m(x)
        return y, ops.LAUNCHES - n0

    ref, n_plain = run("0")
    got, n_default = run("1")
    assert torch.equal(got, ref) and n_default == n_plain - 12, (n_default, n_plain)   # 4 dim-64 + 8 dim-128 (A-stationary) qkv projections
    got, n_all = run("all")
    assert torch.equal(got, ref) and n_all == n_plain - 28, (n_all, n_plain)
