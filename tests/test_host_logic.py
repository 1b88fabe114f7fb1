"""CPU checks of the drop-in boundary's host side: arch registry, state_dict key layout, checkpoint
helpers, weight packing, sharding arithmetic.  No kernel is launched (no GPU needed)."""
import os
import types

import pytest
import torch

from fbanet_b200 import BaseModel, get_arch, load_checkpoint, save_checkpoint
from fbanet_b200.dist import shard_range


def _opt(**kw):
    d = dict(arch="BaseModel", train_ps=40, embed_dim=32, win_size=10, token_projection="linear", token_mlp="leff")
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_get_arch_contract():
    m = get_arch(_opt())
    assert isinstance(m, BaseModel) and m.img_size == 40 and m.embed_dim == 32 and m.window_length == 10
    assert m.num_frames == 14 and m.in_channels == 3  # dataclass defaults of the reference (models/fba_net.py:31-47)
    with pytest.raises(Exception, match="Arch error!"):
        get_arch(_opt(arch="Uformer"))
    with pytest.raises(NotImplementedError):
        get_arch(_opt(token_mlp="ffn"))  # non-default options are out of scope and say so loudly


def test_state_dict_key_layout_matches_appendix_b_and_oracle():
    from oracle.fbanet_oracle import OracleBaseModel
    cfg = dict(num_frames=4, img_size=40, in_channels=3, embed_dim=32, window_length=10)
    m = BaseModel(**cfg, token_projection="linear", token_mlp="leff")
    sd = m.state_dict()
    ref = OracleBaseModel(**cfg).state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in sd)
    for k in ("head.weight", "body.1.body.2.bias", "fusion.feature_fusion.1.weight", "fusion.res_blocks.4.1.body.0.weight",
              "input_proj.proj.1.weight", "HG1_encoderlayer_0.blocks.1.attn.relative_position_bias_table",
              "HG2_decoderlayer_1.blocks.0.attn.qkv.to_kv.weight", "conv_HG1.blocks.0.mlp.dwconv.0.weight",
              "HG1_downsample_0.conv.0.weight", "HG2_upsample_1.deconv.0.bias", "output_proj_HG2_0.proj.0.weight", "tail.0.2.weight", "tail.1.bias"):
        assert k in sd, k
    assert sd["fusion.feature_fusion.0.weight"].shape == (32, 4 * 32, 1, 1)
    assert sd["HG1_upsample_0.deconv.0.weight"].shape == (128, 64, 2, 2)  # torch ConvT layout [in,out,kh,kw]
    assert sd["HG1_encoderlayer_0.blocks.0.attn.relative_position_index"].shape == (100, 100)


def test_full_size_parameter_count():
    m = BaseModel(num_frames=14, img_size=160, in_channels=3, embed_dim=64, window_length=10, token_projection="linear", token_mlp="leff")
    assert sum(p.numel() for p in m.parameters()) == 19_217_237


def test_checkpoint_roundtrip_with_module_prefix(tmp_path):
    cfg = dict(num_frames=4, img_size=40, in_channels=3, embed_dim=32, window_length=10, token_projection="linear", token_mlp="leff")
    a, b = BaseModel(**cfg, seed=1), BaseModel(**cfg, seed=2)
    assert not torch.equal(a.head.weight, b.head.weight)
    state = {"epoch": 7, "state_dict": {"module." + k: v for k, v in a.state_dict().items()}, "optimizer": {}}
    path = save_checkpoint(str(tmp_path), state, "sess")
    assert os.path.basename(path) == "model_epoch_7_sess.pth"
    load_checkpoint(b, path)
    assert all(torch.equal(a.state_dict()[k], b.state_dict()[k]) for k in a.state_dict())
    from fbanet_b200 import load_start_epoch
    assert load_start_epoch(path) == 7
    bad = dict(state, state_dict={"module.nope": torch.zeros(1)})
    torch.save(bad, path)
    with pytest.raises(RuntimeError):
        load_checkpoint(b, path)  # missing / unexpected keys are reported, not swallowed


def test_weight_packing_shapes_and_cache():
    cfg = dict(num_frames=4, img_size=40, in_channels=3, embed_dim=64, window_length=10, token_projection="linear", token_mlp="leff")
    m = BaseModel(**cfg, dtype="bf16")
    P = m.packed()
    assert P["body.0.0.w"].shape == (64, 9 * 64) and P["body.0.0.w"].dtype == torch.bfloat16
    assert P["head.wkc"].shape == (27, 64) and P["head.wkc"].dtype == torch.float32
    assert P["fusion.wsum"].shape == (9, 64)
    assert P["fusion.up0.w"].shape == (4 * 128, 256) and P["fusion.up0.b"].shape == (512,)
    assert P["HG1_encoderlayer_1.0.qkv.w"].shape == (3 * 128, 128)
    assert P["tail.1.w"].shape == (16, 9 * 64)  # final conv padded to the tensor-core minimum N
    # conv weight packing is tap-major / channel-minor: k = (ky*3+kx)*Cin + c
    w = m.body[0].body[0].weight
    assert torch.equal(P["body.0.0.w"][5, (1 * 3 + 2) * 64 + 7].float(), w[5, 7, 1, 2].to(torch.bfloat16).float())
    # PixelShuffle folded into the store: row (2i+j)*E + c  <-  out channel 4c+2i+j
    wt = m.tail[0][0].weight
    assert torch.equal(P["tail.0.0.w"][(2 * 1 + 0) * 64 + 3, 0].float(), wt[4 * 3 + 2, 0, 0, 0].to(torch.bfloat16).float())
    assert m.packed() is P
    with torch.no_grad():
        m.head.bias.add_(1.0)
    assert m.packed() is not P  # in-place parameter change invalidates the cache
    f = BaseModel(**cfg, dtype="fp32").packed()
    assert f["head.w"].shape == (64, 9 * 4) and f["head.w"].dtype == torch.float32


def test_cpu_tensor_is_refused():
    m = BaseModel(num_frames=4, img_size=40, embed_dim=32, window_length=10, token_projection="linear", token_mlp="leff")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 3, 40, 40))


@pytest.mark.parametrize("n,world", [(64, 8), (336, 8), (10, 4), (3, 8), (0, 2)])
def test_shard_range_tiles_exactly(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("H,world", [(1080, 8), (1080, 2), (50, 3), (7, 7)])
def test_band_rows_and_halo_sources(H, world):
    """cfg4 row bands: boundaries tile [0, H) exactly; the rows a rank gathers for its tile rows come from its own band and the
    neighbours the 40-pixel halo (or a reflection at the image border) reaches."""
    from fbanet_b200.dist import band_rows, halo_sources
    r = band_rows(H, world)
    assert r[0] == 0 and r[-1] == H and len(r) == world + 1 and all(b > a for a, b in zip(r, r[1:]))
    assert max(b - a for a, b in zip(r, r[1:])) - min(b - a for a, b in zip(r, r[1:])) <= 1
    if H == 1080 and world == 8:
        assert halo_sources(H, 80, 40, (0, 2), r) == {0: 135, 1: 65}     # rows 0..199 (the top halo reflects into rows 1..40)
        assert halo_sources(H, 80, 40, (12, 14), r) == {7: 135, 6: 25}   # rows 920..1079; the padded rows reflect back inside
    tot = halo_sources(H, min(80, H), min(40, H - 1), (0, -(-H // min(80, H))), r)
    assert sum(tot.values()) == H                                        # all tile rows together touch every image row once


@pytest.mark.skipif(not os.path.exists("/root/reference/fba_net/utils/model_utils.py"), reason="reference tree only exists in the build container")
def test_reference_checkpoint_helpers_accept_our_model(tmp_path):
    """Drop-in boundary exercised with the reference's OWN code: `fba_net/utils/model_utils.py` (torch-only) is loaded
    by file path and its `save_checkpoint` / `load_checkpoint` / `load_checkpoint_multigpu` / `load_start_epoch` / `load_optim`
    are run against our `BaseModel` -- including the `module.`-prefixed keys a `DataParallel` run saves -- and against files
    written by our helpers, and vice versa."""
    import importlib.util
    import torch
    from fbanet_b200 import BaseModel
    from fbanet_b200.utils import model_utils as ours
    spec = importlib.util.spec_from_file_location("ref_model_utils", "/root/reference/fba_net/utils/model_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cfg = dict(num_frames=4, img_size=40, in_channels=3, embed_dim=32, window_length=10, token_projection="linear", token_mlp="leff")
    a, b = BaseModel(**cfg, seed=1), BaseModel(**cfg, seed=2)
    opt = torch.optim.AdamW(a.parameters(), lr=3e-4)
    state = {"epoch": 7, "state_dict": a.state_dict(), "optimizer": opt.state_dict()}
    ref.save_checkpoint(str(tmp_path), state, "ref")                       # the reference writes ...
    path = os.path.join(str(tmp_path), "model_epoch_7_ref.pth")
    ours.load_checkpoint(b, path)                                          # ... we read
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    assert ours.load_start_epoch(path) == ref.load_start_epoch(path) == 7
    assert ours.load_optim(torch.optim.AdamW(b.parameters(), lr=1.0), path) == ref.load_optim(torch.optim.AdamW(b.parameters(), lr=1.0), path) == 3e-4
    c = BaseModel(**cfg, seed=3)
    ours.save_checkpoint(str(tmp_path), {"epoch": 8, "state_dict": {"module." + k: v for k, v in a.state_dict().items()}, "optimizer": opt.state_dict()}, "ours")
    path2 = os.path.join(str(tmp_path), "model_epoch_8_ours.pth")
    ref.load_checkpoint(c, path2)                                          # we write (DataParallel-style keys), the reference's loader reads
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), c.state_dict().values()))
    d = BaseModel(**cfg, seed=4)
    ref.load_checkpoint_multigpu(d, path2)
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), d.state_dict().values()))
    ref.freeze(d)
    assert ref.is_frozen(d) and ours.is_frozen(d)
    ours.unfreeze(d)
    assert not ref.is_frozen(d)


def test_lr_schedules_match_the_reference_scheduler():
    """`train.warmup_cosine_lr` / `step_lr` against the reference's own `GradualWarmupScheduler` + torch schedulers executed over whole
    runs as train.py.bak:104-115,220 drives them (tests/golden/make_golden_reference.py -> lr_schedule_reference.npz)."""
    import os

    import numpy as np
    from fbanet_b200.train import step_lr, warmup_cosine_lr

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lr_schedule_reference.npz"))
    for tag in ("warmup_default", "warmup_short"):
        nepoch, warm, lr0 = int(z[tag + "_cfg"][0]), int(z[tag + "_cfg"][1]), float(z[tag + "_cfg"][2])
        mine = np.array([warmup_cosine_lr(e, lr0, nepoch, warm) for e in range(1, nepoch + 1)])
        np.testing.assert_allclose(mine, z[tag], rtol=1e-9, atol=0)
        assert mine[warm] > lr0 and abs(mine[warm + 1] - lr0) < 1e-12 * lr0  # the hand-over overshoot, then exactly lr
    mine = np.array([step_lr(e) for e in range(1, 251)])
    np.testing.assert_allclose(mine, z["steplr"], rtol=1e-12, atol=0)


def test_drop_path_rates_match_the_reference_constructor():
    """`train.drop_path_rates` against the rates the reference's own constructor handed to its layers (model_structure_reference.json)."""
    import json
    import os

    from fbanet_b200.train import drop_path_rates

    with open(os.path.join(os.path.dirname(__file__), "golden", "model_structure_reference.json")) as f:
        mods = json.load(f)["cfg2_rgb160"]["modules"]
    rates = drop_path_rates()
    for hg in ("HG1", "HG2"):
        for mine, name in (("encoderlayer_0", f"{hg}_encoderlayer_0"), ("encoderlayer_1", f"{hg}_encoderlayer_1"), ("conv", f"conv_{hg}"),
                           ("decoderlayer_0", f"{hg}_decoderlayer_0"), ("decoderlayer_1", f"{hg}_decoderlayer_1")):
            want = [layer["drop_path_rate"] for layer in mods[name]["layers"]]
            assert len(want) == len(rates[mine]) and all(abs(a - b) < 1e-6 for a, b in zip(rates[mine], want)), (name, rates[mine], want)


def test_drop_path_scales_follow_the_reference_layer():
    """layers/drop_path.py:52-63: bernoulli(keep) / keep per burst; p == 0 returns x itself (all ones); keep == 0 skips the division."""
    import torch
    from fbanet_b200 import train
    assert torch.equal(train.drop_path_scales(4, 0.0), torch.ones(4))
    assert torch.equal(train.drop_path_scales(4, 1.0), torch.zeros(4))
    s = train.drop_path_scales(20000, 0.1, generator=torch.Generator().manual_seed(0))
    assert set(s.unique().tolist()) == {0.0, float(torch.tensor(1.0) / 0.9)}
    assert abs(s.mean().item() - 1.0) < 0.01                      # E[noise / keep] = 1: the layer is unbiased
    a = train.drop_path_scales(64, 0.5, generator=torch.Generator().manual_seed(3))
    b = train.drop_path_scales(64, 0.5, generator=torch.Generator().manual_seed(3))
    assert torch.equal(a, b)


# ---- data / weight gradients of the resampling layers as compositions of the FORWARD kernels (train.dgrad_weight_*) ----------------
# The kernels are emulated on the CPU by their documented contracts (include/fbanet_b200.h: packed weight [Cout, taps*Cin] with
# K index tap*Cin + ci, FBANET_STORE_CONVT2 col = (2i+j)*Co + co -> out(2y+i, 2x+j, co), space-to-depth channel (ys*2+xs)*C + c);
# the emulation is first validated against the packings model.py feeds the real kernels for the forward (those are GPU-verified),
# then the gradient packings are checked against autograd.
def _emu_conv_gemm(src, packed, k=1, pad=0, convt2=False):
    import torch
    import torch.nn.functional as F
    Cout, C = packed.shape[0], src.shape[-1]
    w = packed.reshape(Cout, k, k, C).permute(0, 3, 1, 2)
    g = F.conv2d(src.permute(0, 3, 1, 2), w, None, padding=pad).permute(0, 2, 3, 1)          # [N,H,W,Cout]
    if not convt2:
        return g
    N, H, W, _ = g.shape
    Co = Cout // 4
    return g.reshape(N, H, W, 2, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * H, 2 * W, Co)


def _emu_s2d(x):
    N, H, W, C = x.shape
    return x.reshape(N, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H // 2, W // 2, 4 * C)


def test_resampling_layer_gradients_through_forward_kernels():
    import torch
    import torch.nn.functional as F
    from fbanet_b200 import train
    from fbanet_b200.model import BaseModel
    torch.manual_seed(0)
    m = BaseModel(img_size=40, embed_dim=16, window_length=10, token_mlp="leff", dtype="fp32", num_frames=3).double()
    m.compute_dtype = torch.float64
    P = {k: v.double() for k, v in m._pack().items()}
    dd = torch.float64
    # -- the emulation reproduces the forward of the three layers from model.py's own packings
    up, down, ps = m.fusion.upsample0, m.fusion.downsample0, m.tail[0][0]
    x = torch.randn(2, 6, 8, up.weight.shape[0], dtype=dd)
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2), up.weight, None, stride=2).permute(0, 2, 3, 1)
    assert torch.allclose(_emu_conv_gemm(x, P["fusion.up0.w"], convt2=True), want, atol=1e-12)
    xp = torch.randn(2, 6, 8, ps.weight.shape[1], dtype=dd)
    want = F.pixel_shuffle(F.conv2d(xp.permute(0, 3, 1, 2), ps.weight, None, padding=1), 2).permute(0, 2, 3, 1)
    assert torch.allclose(_emu_conv_gemm(xp, P["tail.0.0.w"], 3, 1, convt2=True), want, atol=1e-12)
    assert torch.equal(_emu_s2d(torch.arange(2 * 4 * 6 * 3, dtype=dd).reshape(2, 4, 6, 3))[1, 1, 2, 2 * 3 + 1],
                       torch.arange(2 * 4 * 6 * 3, dtype=dd).reshape(2, 4, 6, 3)[1, 3, 4, 1])      # channel (ys*2+xs)*C + c, ys=1, xs=0

    # -- ConvTranspose2d(2,2): data gradient = 1x1 GEMM over s2d(dy); weight gradient layout
    xr = x.clone().requires_grad_(True)
    wr = up.weight.detach().clone().requires_grad_(True)
    br = torch.zeros(wr.shape[1], dtype=dd, requires_grad=True)
    y = F.conv_transpose2d(xr.permute(0, 3, 1, 2), wr, br, stride=2).permute(0, 2, 3, 1)
    dy = torch.randn_like(y)
    y.backward(dy)
    assert torch.allclose(_emu_conv_gemm(_emu_s2d(dy), train.dgrad_weight_convT2(wr, dd)), xr.grad, atol=1e-12)
    dys = _emu_s2d(dy)                                                                           # what ops.conv_wgrad(x, s2d(dy)) returns:
    dw4 = torch.einsum("nhwo,nhwi->oi", dys, x)[:, :, None, None]
    dw, db = train.convT2_wgrad_layout(dw4, dys.sum((0, 1, 2)), wr.shape[1])
    assert torch.allclose(dw, wr.grad, atol=1e-10) and torch.allclose(db, br.grad, atol=1e-10)

    # -- Conv2d(4, stride 2, pad 1): data gradient = 3x3 GEMM over dy with phase rows + ConvT-style scatter
    xd = torch.randn(2, 8, 12, down.weight.shape[1], dtype=dd, requires_grad=True)
    y = F.conv2d(xd.permute(0, 3, 1, 2), down.weight, None, stride=2, padding=1).permute(0, 2, 3, 1)
    dy = torch.randn_like(y)
    y.backward(dy)
    got = _emu_conv_gemm(dy, train.dgrad_weight_down4(down.weight, dd), 3, 1, convt2=True)
    assert got.shape == xd.shape and torch.allclose(got, xd.grad, atol=1e-12)

    # -- conv3x3 + PixelShuffle(2): data gradient = flipped-weight 3x3 GEMM over s2d(dy) with permuted K; weight gradient layout
    xr = xp.clone().requires_grad_(True)
    wr = ps.weight.detach().clone().requires_grad_(True)
    br = torch.zeros(wr.shape[0], dtype=dd, requires_grad=True)
    y = F.pixel_shuffle(F.conv2d(xr.permute(0, 3, 1, 2), wr, br, padding=1), 2).permute(0, 2, 3, 1)
    dy = torch.randn_like(y)
    y.backward(dy)
    assert torch.allclose(_emu_conv_gemm(_emu_s2d(dy), train.dgrad_weight_pixel_shuffle(wr, dd), 3, 1), xr.grad, atol=1e-12)
    dz = _emu_s2d(dy).permute(0, 3, 1, 2)                                                        # ops.conv_wgrad(x, s2d(dy), 3, 3, 1, 1):
    wz = torch.zeros_like(wr, requires_grad=True)
    bz = torch.zeros_like(br, requires_grad=True)
    F.conv2d(xp.permute(0, 3, 1, 2), wz, bz, padding=1).backward(dz)
    dw, db = train.pixel_shuffle_wgrad_layout(wz.grad, bz.grad)
    assert torch.allclose(dw, wr.grad, atol=1e-10) and torch.allclose(db, br.grad, atol=1e-10)


# ---- the LeWin block's training forward / backward composition (train.lewin_forward_train / lewin_backward) on the CPU -----------
# Every C-ABI op the composition calls is replaced by a torch stand-in with the SAME signature and contract (each real op is pinned
# to exactly that contract by its own -m gpu test); what this checks is the composition: which activation feeds which brick, the
# residual / DropPath gradient sums, where each parameter gradient lands.  Reference: autograd through the oracle's LeWinLayer.
def _install_op_standins(monkeypatch):
    import torch
    import torch.nn.functional as F
    from fbanet_b200 import ops, _lib as L
    from test_gpu_train_spatial import _attention_reference

    def act_fn(act):
        return {L.ACT_NONE: lambda v: v, L.ACT_RELU: F.relu, L.ACT_GELU_TANH: lambda v: F.gelu(v, approximate="tanh"), L.ACT_GELU_ERF: F.gelu}[act]

    def grad_of(fn, inputs, dy):
        leaves = [t.detach().clone().requires_grad_(True) for t in inputs]
        fn(*leaves).backward(dy)
        return [t.grad for t in leaves]

    def conv_gemm(srcs, weight, out, *, bias=None, act=L.ACT_NONE, **kw):
        assert len(srcs) == 1 and not kw and out.shape == (*srcs[0].shape[:3], weight.shape[0])
        y = srcs[0] @ weight.t()
        out.copy_(act_fn(act)(y + bias if bias is not None else y))
        return out

    def conv_wgrad(x, dy, kh=1, kw=1, stride=1, pad=0):
        assert kh == 1 and kw == 1
        return torch.einsum("nhwo,nhwi->oi", dy, x)[:, :, None, None].contiguous(), dy.sum((0, 1, 2))

    def window_attention(qkv, table, B, H, W, heads, win, shift, scale):
        return _attention_reference(qkv, table, B, H, W, qkv.shape[1] // 3, heads, win, shift, scale)

    def window_attention_backward(qkv, dout, table, B, H, W, heads, win, shift, scale):
        C = qkv.shape[1] // 3
        return tuple(grad_of(lambda q, t: _attention_reference(q, t, B, H, W, C, heads, win, shift, scale), [qkv, table], dout))

    def dwconv3x3(x, w9c, b, act):
        C = x.shape[-1]
        return act_fn(act)(F.conv2d(x.permute(0, 3, 1, 2), w9c.t().reshape(C, 1, 3, 3), b, padding=1, groups=C)).permute(0, 2, 3, 1).contiguous()

    def dwconv3x3_backward(x, dy, w9c):
        C = x.shape[-1]
        gx, gw, gb = grad_of(lambda a, w, b: F.conv2d(a.permute(0, 3, 1, 2), w, b, padding=1, groups=C).permute(0, 2, 3, 1),
                             [x, w9c.t().reshape(C, 1, 3, 3), torch.zeros(C, dtype=x.dtype)], dy)
        return gx, gw, gb

    def layernorm_backward(x, dy, gamma, eps=1e-5):
        return tuple(grad_of(lambda a, g, b: F.layer_norm(a, (a.shape[1],), g, b, eps), [x, gamma, torch.zeros_like(gamma)], dy))

    def drop_path_add(x, scale, skip=None):
        v = x * scale.view(-1, *([1] * (x.dim() - 1)))
        return v if skip is None else skip + v

    monkeypatch.setattr(ops, "conv_gemm", conv_gemm)
    monkeypatch.setattr(ops, "conv_wgrad", conv_wgrad)
    monkeypatch.setattr(ops, "layernorm", lambda x, g, b, eps=1e-5: F.layer_norm(x, (x.shape[1],), g, b, eps))
    monkeypatch.setattr(ops, "layernorm_backward", layernorm_backward)
    monkeypatch.setattr(ops, "window_attention", window_attention)
    monkeypatch.setattr(ops, "window_attention_backward", window_attention_backward)
    monkeypatch.setattr(ops, "dwconv3x3", dwconv3x3)
    monkeypatch.setattr(ops, "dwconv3x3_backward", dwconv3x3_backward)
    def act_forward(x, act, alpha=None):
        return F.prelu(x, alpha) if act == L.ACT_PRELU else act_fn(act)(x)

    def act_backward(x, dy, act, alpha=None):
        if act == L.ACT_PRELU:
            return tuple(grad_of(F.prelu, [x, alpha], dy))
        return grad_of(act_fn(act), [x], dy)[0]

    def gate_forward(feat, wsum):                                            # feat [B,F,H,W,C], wsum [9,C]: the collapsed gate
        B, Fr, H, W, C = feat.shape
        w = wsum.t().reshape(1, C, 3, 3).to(feat.dtype).contiguous()
        sc = F.conv2d(feat.reshape(B * Fr, H, W, C).permute(0, 3, 1, 2), w, None, padding=1).view(B, Fr, H, W)
        gate = torch.sigmoid((sc[:, 1:] - sc[:, :1]).abs())
        gated = torch.cat([feat[:, :1], feat[:, 1:] * gate[..., None]], 1)
        return gate, gated.permute(0, 2, 3, 1, 4).reshape(B, H, W, Fr * C)

    def faf_gate(feat, wsum, want_gate=True, want_gated=False, score=None):
        gate, gated = gate_forward(feat, wsum)
        return (gate, gated) if (want_gate and want_gated) else (gated if want_gated else gate)

    def faf_gate_backward(feat, dgated, gate, score, wsum):
        gf, gw = grad_of(lambda f, w: gate_forward(f, w)[1], [feat, wsum.to(feat.dtype)], dgated)
        return gf, gw

    monkeypatch.setattr(ops, "act_forward", act_forward)
    monkeypatch.setattr(ops, "act_backward", act_backward)
    monkeypatch.setattr(ops, "faf_gate", faf_gate)
    monkeypatch.setattr(ops, "faf_gate_backward", faf_gate_backward)
    monkeypatch.setattr(ops, "drop_path_add", drop_path_add)


@pytest.mark.parametrize("shift", [0, 2])
def test_lewin_training_step_composition_matches_autograd_of_the_oracle_layer(monkeypatch, shift):
    import torch
    from fbanet_b200 import train
    from fbanet_b200.model import _Layer
    from oracle.fbanet_oracle import LeWinLayer
    _install_op_standins(monkeypatch)
    torch.manual_seed(3)
    dim, res, heads, win, B = 16, (8, 8), 2, 4, 3
    ly = _Layer(dim, res, heads, win, shift, 4.0).double()
    with torch.no_grad():
        for p in ly.parameters():
            p.copy_(torch.randn_like(p) * 0.3)
    ref = LeWinLayer(dim, res, heads, win, shift, 4.0, "tanh").double()
    ref.load_state_dict(ly.state_dict())
    x = torch.randn(B, *res, dim, dtype=torch.float64)
    dy = torch.randn_like(x)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr.view(B, -1, dim)).view_as(x)
    yr.backward(dy)
    y, saved = train.lewin_forward_train(ly, x)
    assert torch.allclose(y, yr.detach(), atol=1e-10)
    dx = train.lewin_backward(ly, saved, dy)
    assert torch.allclose(dx, xr.grad, atol=1e-9)
    names = dict(ly.named_parameters())
    assert len(names) == 17
    for n, pr in ref.named_parameters():
        assert names[n].grad is not None and torch.allclose(names[n].grad, pr.grad, atol=1e-8), n
    # gradients ACCUMULATE (flat gradient buffer semantics): a second backward doubles them
    train.lewin_backward(ly, saved, dy)
    for n, pr in ref.named_parameters():
        assert torch.allclose(names[n].grad, 2 * pr.grad, atol=1e-8), n
    # DropPath: a burst whose two branches are dropped passes x and dy through untouched and contributes nothing to the parameters
    for p in ly.parameters():
        p.grad = None
    s = torch.tensor([1.0, 0.0, 1.0], dtype=torch.float64)
    y2, saved2 = train.lewin_forward_train(ly, x, s_attn=s / 0.75, s_mlp=s / 0.75)
    dx2 = train.lewin_backward(ly, saved2, dy)
    assert torch.equal(y2[1], x[1]) and torch.equal(dx2[1], dy[1])
    keep = [0, 2]
    for p in ly.parameters():
        p.grad, p.gsave = None, p.grad.clone()
    y3, saved3 = train.lewin_forward_train(ly, x[keep].contiguous(), s_attn=s[keep] / 0.75, s_mlp=s[keep] / 0.75)
    train.lewin_backward(ly, saved3, dy[keep].contiguous())
    assert torch.allclose(y3, y2[keep], atol=1e-12)
    for n, p in ly.named_parameters():
        assert torch.allclose(p.grad, p.gsave, atol=1e-9), n


def _install_conv_standins(monkeypatch):
    """General conv stand-ins on top of :func:`_install_op_standins`: the implicit GEMM by its packed-weight contract (any k / stride /
    pad, ConvT-style scatter store), space-to-depth, and the weight gradient via autograd of the same conv."""
    import torch
    import torch.nn.functional as F
    from fbanet_b200 import ops, _lib as L

    def conv_gemm(srcs, weight, out, *, kh=1, kw=1, stride=1, pad=0, bias=None, act=L.ACT_NONE, store_mode=L.STORE_NHWC, base=None,
                  cout_store=None):
        assert len(srcs) == 1 and act == L.ACT_NONE
        src = srcs[0]
        w = weight.reshape(weight.shape[0], kh, kw, src.shape[-1]).permute(0, 3, 1, 2)
        g = F.conv2d(src.permute(0, 3, 1, 2), w, bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
        if store_mode == L.STORE_NCHW_BASE:      # planar out + bilinear x4 of the base frame (half-pixel centres, edge clamp)
            assert cout_store == weight.shape[0] and base.shape[1] == cout_store
            g = g.permute(0, 3, 1, 2) + F.interpolate(base.to(g.dtype), scale_factor=4, mode="bilinear", align_corners=False)
        if store_mode == L.STORE_CONVT2:
            N, H, W, C4 = g.shape
            g = g.reshape(N, H, W, 2, 2, C4 // 4).permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * H, 2 * W, C4 // 4)
        assert out.shape == g.shape, (out.shape, g.shape)
        out.copy_(g)
        return out

    def conv_wgrad(x, dy, kh=1, kw=1, stride=1, pad=0):
        w = torch.zeros(dy.shape[-1], x.shape[-1], kh, kw, dtype=x.dtype, requires_grad=True)
        b = torch.zeros(dy.shape[-1], dtype=x.dtype, requires_grad=True)
        F.conv2d(x.permute(0, 3, 1, 2), w, b, stride=stride, padding=pad).backward(dy.permute(0, 3, 1, 2))
        return w.grad, b.grad

    monkeypatch.setattr(ops, "conv_gemm", conv_gemm)
    monkeypatch.setattr(ops, "conv_wgrad", conv_wgrad)
    monkeypatch.setattr(ops, "space_to_depth", lambda x: _emu_s2d(x).contiguous())

    def to_nhwc(x, cp, dtype):                   # planar [N,C,H,W] -> channels-last, channels zero-padded to cp
        return F.pad(x.permute(0, 2, 3, 1), (0, cp - x.shape[1])).to(dtype).contiguous()

    def training_loss(restored, target, clamp_restored=False):
        from oracle.fbanet_oracle import charbonnier_loss, gw_loss
        r = restored.detach().to(target.dtype).requires_grad_(True)
        rc = torch.clamp(r, 0.0, 1.0) if clamp_restored else r
        c, g = charbonnier_loss(rc, target), gw_loss(rc, target)
        (c + 3.0 * g).backward()
        return torch.stack([c + 3.0 * g, c, g]).detach().double(), r.grad        # (total, charbonnier, gw) like the kernel

    monkeypatch.setattr(ops, "to_nhwc", to_nhwc)
    monkeypatch.setattr(ops, "training_loss", training_loss)


def test_hourglass_training_composition_matches_autograd_of_the_oracle(monkeypatch):
    """train.hourglass_forward_train + Tape.backward (ten LeWin layers, two 4x4 s2 downsamples, two transposed-conv upsamples, two
    skip concats whose sources receive summed gradients) against autograd through the oracle's first hourglass
    (models/fba_net.py:271-287), op stand-ins as above; then DropPath in training mode: reproducible from the generator, and a rate
    of zero everywhere reproduces the eval forward."""
    import torch
    from fbanet_b200 import train
    from fbanet_b200.model import BaseModel
    from oracle.fbanet_oracle import OracleBaseModel
    _install_op_standins(monkeypatch)
    _install_conv_standins(monkeypatch)
    cfg = dict(num_frames=2, img_size=16, embed_dim=16, window_length=4)
    m = BaseModel(token_mlp="leff", dtype="fp32", seed=1, **cfg).double()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "relative_position_bias_table" in n or (p.dim() == 1 and "norm" not in n):
                p.copy_(torch.randn_like(p) * 0.2)                          # biases / tables away from their zero-ish init
    o = OracleBaseModel(**cfg).double()
    missing = o.load_state_dict(m.state_dict(), strict=False)
    assert not missing.missing_keys, missing.missing_keys
    B, S, E = 2, 16, 16
    y = torch.randn(B, S, S, E, dtype=torch.float64)
    dout = torch.randn(B, S, S, 2 * E, dtype=torch.float64)
    yr = y.clone().requires_grad_(True)
    ref, _ = o._hourglass("HG1", yr.view(B, S * S, E))
    ref.backward(dout.view(B, S * S, 2 * E))
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    out, tape = train.hourglass_forward_train(m, "HG1", y, training=False)
    assert torch.allclose(out.view(B, S * S, 2 * E), ref.detach(), atol=1e-9)
    grads = tape.backward(out, dout)
    assert set(grads) == {id(y)} and torch.allclose(grads[id(y)], yr.grad, atol=1e-8)
    got, n_checked = dict(m.named_parameters()), 0
    for n, pr in o.named_parameters():
        if pr.grad is None:
            assert got[n].grad is None, n                                 # nothing outside the hourglass was touched
            continue
        assert got[n].grad is not None and torch.allclose(got[n].grad, pr.grad, atol=1e-7), n
        n_checked += 1
    assert n_checked == 10 * 17 + 4 * 2                                   # ten LeWin layers, two downsamples, two upsamples
    # training mode: per-layer rates linspace(0, 0.1, 8) -> the first layer never drops, the rest draw from the generator
    a, _ = train.hourglass_forward_train(m, "HG1", y, generator=torch.Generator().manual_seed(7))
    b, _ = train.hourglass_forward_train(m, "HG1", y, generator=torch.Generator().manual_seed(7))
    assert torch.equal(a, b) and not torch.allclose(a, out)
    m.drop_path_rate = 0.0
    c, _ = train.hourglass_forward_train(m, "HG1", y, generator=torch.Generator().manual_seed(7))
    assert torch.equal(c, out)


def test_faf_block_training_composition_matches_autograd_of_the_oracle(monkeypatch):
    """train.faf_forward_train + Tape.backward: the whole FAFBlock (gate, 1x1 fusion + PReLU(0.1), ten ResBlocks, 4x4 s2 / transposed
    conv resampling, fusion_tail, skip) against autograd through the oracle's FAFBlock AS WRITTEN (two embedding convs): output,
    d feat, every parameter gradient -- temporal_attn0 and both embedding biases must come out (numerically) zero in the reference
    and untouched in the composition."""
    import torch
    from fbanet_b200 import train
    from fbanet_b200.model import _FAF
    from oracle.fbanet_oracle import FAFBlock
    _install_op_standins(monkeypatch)
    _install_conv_standins(monkeypatch)
    torch.manual_seed(11)
    E, Fr, B, S = 8, 3, 2, 8
    fu = _FAF(E, Fr).double()
    with torch.no_grad():
        for p in fu.parameters():
            p.copy_(torch.randn_like(p) * (0.3 if p.dim() == 1 else 0.7 / p[0].numel() ** 0.5))
    ref = FAFBlock(E, Fr).double()
    ref.load_state_dict(fu.state_dict())
    feat = torch.randn(B, Fr, S, S, E, dtype=torch.float64)
    dout = torch.randn(B, S, S, E, dtype=torch.float64)
    fr = feat.clone().requires_grad_(True)
    yr = ref(fr.permute(0, 1, 4, 2, 3))                                      # [B,E,S,S]
    yr.backward(dout.permute(0, 3, 1, 2))
    # the gate kernel's contract takes wsum in fp32, so even this float64 run carries one 2^-24 rounding: bounds are 1e-6 relative
    def close(a, b):
        return (a - b).abs().max().item() <= 1e-6 * b.abs().max().item()
    out, tape = train.faf_forward_train(fu, feat)
    assert close(out, yr.detach().permute(0, 2, 3, 1))
    grads = tape.backward(out, dout)
    assert set(grads) == {id(feat)} and close(grads[id(feat)], fr.grad)
    got = dict(fu.named_parameters())
    for n, pr in ref.named_parameters():
        if n.startswith("temporal_attn0") or n == "temporal_attn1.bias":
            assert pr.grad.abs().max().item() < 1e-12 and got[n].grad is None, n
            continue
        assert got[n].grad is not None and close(got[n].grad, pr.grad), n


def test_whole_model_training_backward_matches_autograd_of_the_oracle(monkeypatch):
    """train.model_forward_train + the training loss + Tape.backward: the WHOLE model (head, body, FAF block, input projection, both
    hourglasses incl. HG2's four-way projections, output projections, two PixelShuffle stages, final conv + bilinear base) against
    autograd through the oracle model and the oracle's CharbonnierLoss + 3 GWLoss: restored image, loss, and the gradient of every
    parameter (models/fba_net.py:242-322, train.py.bak:163-169).  Op stand-ins as above."""
    import torch
    from fbanet_b200 import train, ops
    from fbanet_b200.model import BaseModel
    from oracle.fbanet_oracle import OracleBaseModel, training_loss
    _install_op_standins(monkeypatch)
    _install_conv_standins(monkeypatch)
    cfg = dict(num_frames=2, img_size=16, embed_dim=16, window_length=4)
    m = BaseModel(token_mlp="leff", dtype="fp32", seed=2, **cfg).double()
    m.compute_dtype = torch.float64
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "relative_position_bias_table" in n or (p.dim() == 1 and "norm" not in n):
                p.copy_(torch.randn_like(p) * 0.2)
    o = OracleBaseModel(**cfg).double()
    o.load_state_dict(m.state_dict())
    B = 2
    burst = torch.rand(B, 2, 3, 16, 16, dtype=torch.float64)
    target = torch.rand(B, 3, 64, 64, dtype=torch.float64)
    ref = o(burst)
    loss_ref = training_loss(ref, target, clamp_restored=True)             # train.py.bak:167-168
    loss_ref.backward()
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    monkeypatch.setattr(torch.Tensor, "float", lambda self: self)           # keep float64 through the composition's fp32 casts

    # two fp32 roundings survive in this float64 run, both part of the kernels' contracts: wsum of the FAF gate and the planar fp32
    # restored image; bounds are 1e-5 relative
    def close(a, b, tol=1e-5):
        return (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-30)
    restored, tape = train.model_forward_train(m, burst, training=False)
    assert restored.dtype == torch.float32 and close(restored.double(), ref.detach(), 1e-6)
    loss, d_restored = ops.training_loss(restored, target, clamp_restored=True)
    assert abs(loss[0].item() - loss_ref.item()) < 1e-7
    left = tape.backward(restored, d_restored)
    assert left == {}                                                        # the burst itself takes no gradient
    got, n_checked = dict(m.named_parameters()), 0
    for n, pr in o.named_parameters():
        if n.startswith("fusion.temporal_attn0") or n == "fusion.temporal_attn1.bias":
            assert got[n].grad is None and pr.grad.abs().max().item() < 1e-12, n
            continue
        assert got[n].grad is not None and close(got[n].grad, pr.grad), (n, (got[n].grad - pr.grad).abs().max().item(), pr.grad.abs().max().item())
        n_checked += 1
    assert n_checked == len(got) - 3


def test_train_step_plumbing(monkeypatch):
    """train.train_step: zero the flat gradient, training-mode forward, loss, tape backward INTO the flat gradient buffer (the
    parameters' .grad are views of it), all-reduce scale handed to the fused AdamW step (train.py.bak:163-170).  Stand-ins as above
    plus the AdamW update formula for fbanet_adam_step_sm100 (the kernel itself is pinned to torch.optim by its -m gpu test)."""
    import torch
    from fbanet_b200 import train, ops
    from fbanet_b200.model import BaseModel
    _install_op_standins(monkeypatch)
    _install_conv_standins(monkeypatch)
    calls = []

    def adam_step(param, grad, m, v, step, lr, betas, eps, wd, decoupled, grad_scale):
        calls.append((step, lr, wd, decoupled, grad_scale))
        g = grad * grad_scale
        param.mul_(1 - lr * wd)
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        param.addcdiv_(m / (1 - betas[0] ** step), (v / (1 - betas[1] ** step)).sqrt() + eps, value=-lr)
    monkeypatch.setattr(ops, "adam_step", adam_step)
    m = BaseModel(token_mlp="leff", dtype="fp32", seed=3, num_frames=2, img_size=16, embed_dim=16, window_length=4)
    for p in m.parameters():
        p.requires_grad_(True)
    m.drop_path_rate = 0.0                                                  # (with one burst a dropped branch zeroes its layer's gradients)
    flat = train.FlatParams(m.parameters())
    before = flat.data.clone()
    burst, target = torch.rand(1, 2, 3, 16, 16), torch.rand(1, 3, 64, 64)
    gen = torch.Generator().manual_seed(5)
    l1 = train.train_step(m, flat, burst, target, lr=1e-3, generator=gen)
    assert calls == [(1, 1e-3, 0.02, True, 1.0)] and torch.isfinite(l1).all() and l1[0].item() > 0
    used = flat.grad != 0
    assert used.float().mean().item() > 0.9                                # all but the cancelled FAF embedding parameters
    names = {id(p): n for n, p in m.named_parameters()}
    dead = [names[id(p)] for p in flat.params if not p.grad.any()]
    assert sorted(dead) == ["fusion.temporal_attn0.bias", "fusion.temporal_attn0.weight", "fusion.temporal_attn1.bias"], dead
    moved = (flat.data - before).abs()
    assert 0.5e-3 < moved[used].median().item() < 1.5e-3                    # first Adam step: |update| ~ lr
    assert all(p.data.data_ptr() == flat.data[o:].data_ptr() for p, o in zip(flat.params, flat.offsets))   # still views
    # the inference path's weight cache keys on FlatParams.generation (the real optimizer kernel bumps no version counter)
    sig = m._signature()
    flat.generation += 1
    assert m._signature() != sig
    l2 = train.train_step(m, flat, burst, target, lr=1e-3, generator=gen)
    assert calls[-1][0] == 2 and l2[0].item() < l1[0].item()                      # the same sample again: the loss went down
    # bucketed reduction driven by the tape: with small buckets, all but the buckets holding the three unreached FAF parameters (and
    # the ones queued behind them: buckets start in order) were complete -- i.e. would have been on the wire -- before finish_reduce
    seen = {}
    orig_finish = train.FlatParams.finish_reduce

    def finish(self):
        seen["launched"], seen["buckets"] = sum(self._launched), len(self._buckets)
        seen["first_blocked"] = self._launched.index(False)
        return orig_finish(self)
    monkeypatch.setattr(train.FlatParams, "finish_reduce", finish)
    train.train_step(m, flat, burst, target, lr=1e-3, generator=gen, bucket_bytes=64 << 10)
    b, e, _ = flat._buckets[seen["first_blocked"]]
    blocked = [names[id(p)] for p, o in zip(flat.params, flat.offsets) if b <= o < e]
    assert seen["buckets"] > 10 and seen["launched"] >= seen["buckets"] - 3 and any(n.startswith("fusion.temporal_attn") for n in blocked), seen
    assert train._REDUCER is None and all(flat._launched)


def test_flat_params_optimizer_state_round_trips_with_torch_adamw(tmp_path):
    """Resume: FlatParams speaks torch.optim's state_dict format (train.py.bak:199-246 saves ``optimizer.state_dict()``,
    utils/model_utils.py:51-62 ``load_optim`` restores it and reads the learning rate from ``param_groups``)."""
    import copy
    import torch
    from fbanet_b200.train import FlatParams
    from fbanet_b200.utils.model_utils import load_optim, save_checkpoint
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.PReLU(), torch.nn.Linear(7, 3))
    twin = copy.deepcopy(net)
    opt = torch.optim.AdamW(twin.parameters(), lr=3e-4, weight_decay=0.02)
    for _ in range(3):
        opt.zero_grad()
        twin(torch.randn(4, 5)).square().mean().backward()
        opt.step()
    path = save_checkpoint(str(tmp_path), {"epoch": 3, "state_dict": twin.state_dict(), "optimizer": opt.state_dict()}, "s")
    flat = FlatParams(net.parameters())
    assert flat.state_dict()["state"] == {} and len(flat.state_dict()["param_groups"][0]["params"]) == 5
    lr = load_optim(flat, path)                                            # the reference helper, FlatParams in the optimizer's place
    assert lr == 3e-4 and flat.step == 3 and flat.param_groups[0]["weight_decay"] == 0.02
    for i, (p, o) in enumerate(zip(flat.params, flat.offsets)):
        st = opt.state[list(twin.parameters())[i]]
        assert torch.equal(flat.exp_avg[o:o + p.numel()].view_as(p), st["exp_avg"])
        assert torch.equal(flat.exp_avg_sq[o:o + p.numel()].view_as(p), st["exp_avg_sq"])
    # and back: a torch optimizer resumes from what FlatParams writes
    opt2 = torch.optim.AdamW(copy.deepcopy(twin).parameters(), lr=1.0)
    opt2.load_state_dict(flat.state_dict())
    assert opt2.param_groups[0]["lr"] == 3e-4
    for a, b in zip(opt.state.values(), opt2.state.values()):
        assert float(a["step"]) == float(b["step"]) and torch.equal(a["exp_avg"], b["exp_avg"]) and torch.equal(a["exp_avg_sq"], b["exp_avg_sq"])


def test_fit_epoch_loop_schedule_checkpoints_and_resume(monkeypatch, tmp_path):
    """train.fit: per-epoch learning rates of the reference's warm-up + cosine schedule reach the optimizer step, the reference's
    checkpoint files appear (train.py.bak:236-245), and a fresh model + FlatParams resumed from model_latest.pth through the
    reference's helpers continues bit-identically to the uninterrupted run (op stand-ins as above)."""
    import torch
    from fbanet_b200 import train, ops
    from fbanet_b200.model import BaseModel
    from fbanet_b200.utils.model_utils import load_checkpoint, load_optim, load_start_epoch
    _install_op_standins(monkeypatch)
    _install_conv_standins(monkeypatch)
    lrs = []

    def adam_step(param, grad, m, v, step, lr, betas, eps, wd, decoupled, grad_scale):
        lrs.append(lr)
        g = grad * grad_scale
        param.mul_(1 - lr * wd)
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        param.addcdiv_(m / (1 - betas[0] ** step), (v / (1 - betas[1] ** step)).sqrt() + eps, value=-lr)
    monkeypatch.setattr(ops, "adam_step", adam_step)

    def make():
        m = BaseModel(token_mlp="leff", dtype="fp32", seed=3, num_frames=2, img_size=16, embed_dim=16, window_length=4)
        m.drop_path_rate = 0.0
        for p in m.parameters():
            p.requires_grad_(True)
        return m, train.FlatParams(m.parameters())
    g = torch.Generator().manual_seed(4)
    batches = [(torch.rand(1, 2, 3, 16, 16, generator=g), torch.rand(1, 3, 64, 64, generator=g)) for _ in range(2)]
    m, flat = make()
    logged = []
    hist = train.fit(m, flat, batches, nepoch=4, lr_initial=1e-3, warmup_epochs=2, model_dir=str(tmp_path), checkpoint_every=2,
                     log=lambda e, lr, l: logged.append((e, lr)))
    want = [train.warmup_cosine_lr(e, 1e-3, 4, 2) for e in (1, 2, 3, 4)]
    assert [lr for _, lr in logged] == want and lrs == [lr for lr in want for _ in batches]
    assert len(hist) == 4 and hist[-1] < hist[0]
    assert sorted(f.name for f in tmp_path.iterdir()) == ["model_epoch_2.pth", "model_epoch_4.pth", "model_latest.pth"]
    final = flat.data.clone()
    # resume after epoch 2 with the reference's helpers
    m2, flat2 = make()
    ck = str(tmp_path / "model_epoch_2.pth")
    load_checkpoint(m2, ck)
    assert load_start_epoch(ck) == 2 and load_optim(flat2, ck) == want[1] and flat2.step == 4
    train.fit(m2, flat2, batches, nepoch=4, start_epoch=3, lr_initial=1e-3, warmup_epochs=2)
    assert torch.equal(flat2.data, final)


def test_torch_library_ops_are_registered_with_fake_kernels_and_no_cpu_fallback():
    """the C-ABI launchers are reachable as ``torch.library`` custom ops (``torch.ops.fbanet.*``); shapes propagate under
    FakeTensorMode; a CPU tensor finds no kernel (there is no CPU fallback)."""
    import pytest
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    import fbanet_b200.torch_ops as T

    for n in T.OPS:
        assert hasattr(torch.ops.fbanet, n), n
    with FakeTensorMode():
        b = torch.empty(2, 14, 3, 40, 40)
        assert torch.ops.fbanet.warp(b, torch.empty(2, 14, 3, 3, dtype=torch.float64)).shape == b.shape
        assert torch.ops.fbanet.forward(b, 0).shape == (2, 3, 160, 160)
        x = torch.empty(2, 40, 40, 64, dtype=torch.bfloat16)
        assert torch.ops.fbanet.conv3x3(x, torch.empty(128, 576, dtype=torch.bfloat16), None, None, 0).shape == (2, 40, 40, 128)
        assert torch.ops.fbanet.linear(x, torch.empty(192, 64, dtype=torch.bfloat16), None, None, 0).shape == (2, 40, 40, 192)
        assert torch.ops.fbanet.layernorm(x, torch.empty(64), torch.empty(64), 1e-5).shape == x.shape
        assert torch.ops.fbanet.window_attention(torch.empty(2, 40, 40, 192, dtype=torch.bfloat16), torch.empty(361, 4), 4, 10, 5, 0.25).shape == x.shape
        f = torch.empty(2, 14, 40, 40, 64, dtype=torch.bfloat16)
        assert torch.ops.fbanet.faf_gate_fuse(f, torch.empty(9, 64), torch.empty(64, 896, dtype=torch.bfloat16), torch.empty(64),
                                              torch.empty(1)).shape == (2, 40, 40, 64)
    with pytest.raises(NotImplementedError):
        torch.ops.fbanet.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8), 1e-5)


def test_round_rowsum_keeps_row_sums_and_touches_few_elements():
    """ops.round_rowsum: bf16 rows whose sums equal the fp32 rows' sums (plain rounding leaves ~sqrt(K) * 2^-10 * |w|), at most nine
    elements per row moved off round-to-nearest, zero (padding) columns untouched, fp32 a plain cast."""
    import torch
    from fbanet_b200 import ops

    g = torch.Generator().manual_seed(3)
    w = (torch.rand(96, 576, generator=g) * 2 - 1) / 24
    w[:, 570:] = 0
    plain = w.to(torch.bfloat16)
    r = ops.round_rowsum(w, torch.bfloat16)
    assert r.dtype == torch.bfloat16 and r.shape == w.shape
    err_plain = (plain.double().sum(1) - w.double().sum(1)).abs()
    err = (r.double().sum(1) - w.double().sum(1)).abs()
    assert err.max() < 2e-6 and err.max() < 0.02 * err_plain.mean(), (err.max(), err_plain.mean())
    assert ((r != plain).sum(1) <= 9).all()
    assert (r[:, 570:] == 0).all()
    assert (r.float() - w).abs().max() < 5e-4            # the moved elements change by about one ulp of a typical weight (max |w| ulp = 2.4e-4)
    assert torch.equal(ops.round_rowsum(w, torch.float32), w)
