"""CPU: the oracle's modules reproduce outputs of the REFERENCE'S OWN layer code (tests/golden/layers_reference.npz).

The fixture was produced in the build container by `tests/golden/make_golden_layers.py`, which imports the unmodified
`fba_net.layers.*` / `fba_net.blocks.*` from /root/reference on top of numpy stand-ins for the jax / equinox primitives
(`tests/golden/jaxshim/`) and executes every layer that can run as written.  Weights are stored as the reference holds them;
the converters below are the layout statements of SURVEY A-11 (the oracle keeps the torch / checkpoint convention):

* `Conv2dLayer` swaps H and W around an Equinox conv (layers/conv2d.py:29-46, swap_channels_last_to_third_from_last.py:13-15), so the
  kernel's first spatial axis runs along W: torch weight = reference weight with its two spatial axes transposed.
* `ConvTranspose2dLayer` (layers/conv2d_transpose.py:9-29) does the same around an Equinox ConvTranspose2d, which stores
  `[out, in, kh, kw]` and does not flip the kernel: torch weight `[in, out, a, b]` = reference weight `[out, in, kw-1-b, kh-1-a]`.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fbanet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "layers_reference.npz")
TOL = dict(rtol=2e-5, atol=2e-5)


@pytest.fixture(scope="module")
def ref():
    z = np.load(GOLD)
    return {k: z[k] for k in z.files}


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def conv_w(w):
    return t(w).transpose(2, 3).contiguous()


def convT_w(w):
    return t(w).flip(2, 3).transpose(2, 3).transpose(0, 1).contiguous()


def bias(b):
    return t(b).reshape(-1)


def to_chw(x):  # [H,W,C] -> [1,C,H,W]
    return t(x).permute(2, 0, 1)[None]


def to_hwc(y):  # [1,C,H,W] -> [H,W,C]
    return y[0].permute(1, 2, 0)


def load_conv(conv, ref, prefix):
    w = ref[prefix + ".weight"]
    conv.weight.data = convT_w(w) if isinstance(conv, torch.nn.ConvTranspose2d) else conv_w(w)
    conv.bias.data = bias(ref[prefix + ".bias"])


def close(a, b):
    np.testing.assert_allclose(a.detach().numpy() if torch.is_tensor(a) else a, b, **TOL)


@pytest.mark.parametrize("case,kw", [("conv3x3", dict(padding=1)), ("conv1x1", dict()), ("conv4x4s2", dict(stride=2, padding=1)),
                                      ("depthwise", dict(padding=1, groups=6))])
def test_reference_layers_conv(ref, case, kw):
    y = F.conv2d(to_chw(ref[f"{case}/x"]), conv_w(ref[f"{case}/conv.weight"]), bias(ref[f"{case}/conv.bias"]), **kw)
    close(to_hwc(y), ref[f"{case}/y"])


def test_reference_layers_conv_needs_the_transposed_kernel(ref):
    """Guards the pin itself: with the weight taken verbatim the non-square 3x3 case must NOT match (A-11 is a real difference)."""
    y = F.conv2d(to_chw(ref["conv3x3/x"]), t(ref["conv3x3/conv.weight"]), bias(ref["conv3x3/conv.bias"]), padding=1)
    assert np.abs(to_hwc(y).numpy() - ref["conv3x3/y"]).max() > 1e-2


def test_reference_layers_conv_transpose_is_oracle_upsample(ref):
    up = O.Upsample(4, 3)
    load_conv(up.deconv[0], ref, "convT2x2s2/conv")
    x = ref["convT2x2s2/x"]
    H, W, C = x.shape
    y = up(t(x).reshape(1, H * W, C), H, W)  # tokens in, tokens out (layers/upsample.py)
    close(y.reshape(2 * H, 2 * W, 3), ref["convT2x2s2/y"])


def test_reference_layers_downsample_is_oracle_downsample(ref):
    dn = O.Downsample(4, 6)
    load_conv(dn.conv[0], ref, "conv4x4s2/conv")
    x = ref["conv4x4s2/x"]
    H, W, C = x.shape
    y = dn(t(x).reshape(1, H * W, C), H, W)
    close(y.reshape(H // 2, W // 2, 6), ref["conv4x4s2/y"])


def test_reference_layers_resblock(ref):
    rb = O.ResBlock(4)
    load_conv(rb.body[0], ref, "resblock/body.0")
    load_conv(rb.body[2], ref, "resblock/body.2")
    close(to_hwc(rb(to_chw(ref["resblock/x"]))), ref["resblock/y"])


def test_reference_layers_input_projection(ref):
    p = O.Proj(4, 6)
    load_conv(p.proj[0], ref, "input_proj/proj.0")
    p.proj[1].weight.data = t(ref["input_proj/proj.1.weight"]).reshape(1)
    assert p.proj[1].weight.item() == 0.25  # A-14: nn.PReLU() default
    y = to_hwc(p(to_chw(ref["input_proj/x"])))
    assert (y < 0).any()
    close(y.reshape(36, 6), ref["input_proj/y"])  # "(height width) channels" token order


@pytest.mark.parametrize("heads", [1, 2, 4])
def test_reference_layers_linear_projection(ref, heads):
    pre = f"linear_projection_h{heads}/"
    lp = O.LinearProjection(8, heads)
    lp.load_state_dict({k: t(ref[pre + k]) for k in ("to_q.weight", "to_q.bias", "to_kv.weight", "to_kv.bias")})
    q, k, v = lp(t(ref[pre + "x"])[None])
    close(q[0], ref[pre + "q"])  # [h, n, d/h]
    close(k[0], ref[pre + "k"])
    close(v[0], ref[pre + "v"])


def swin_table_from_as_written(table, idx_written, w):
    """The reference's index (window_attention.py:70-90) is `(dy + dx + 2w - 2) * (2w - 1)`, out of range for the `(2w-1)^2`-row table
    and clamped by jnp's gather (A-2).  The oracle uses the Swin index; this re-tabulates the bias the reference actually added
    as a function of (dy, dx), so that everything else in the attention can be compared."""
    ys, xs = np.divmod(np.arange(w * w), w)
    dy, dx = ys[:, None] - ys[None, :], xs[:, None] - xs[None, :]
    assert np.array_equal(idx_written, (dy + dx + 2 * w - 2) * (2 * w - 1))
    added = table[np.clip(idx_written, 0, table.shape[0] - 1)]  # [N, N, heads]
    out = np.zeros_like(table)
    out[(dy + w - 1) * (2 * w - 1) + (dx + w - 1)] = added
    return out


def load_attention(attn, ref, pre, w):
    attn.relative_position_bias_table.data = t(swin_table_from_as_written(ref[pre + "relative_position_bias_table"],
                                                                           ref[pre + "relative_position_index_as_written"].astype(np.int64), w))
    for k in ("qkv.to_q.weight", "qkv.to_q.bias", "qkv.to_kv.weight", "qkv.to_kv.bias", "proj.weight", "proj.bias"):
        mod, leaf = k.rsplit(".", 1)
        getattr(attn.get_submodule(mod), leaf).data = t(ref[pre + k])


def test_reference_layers_window_attention(ref):
    attn = O.WindowAttention(8, 3, 1)
    load_attention(attn, ref, "window_attention/", 3)
    close(attn(t(ref["window_attention/x"])[None])[0], ref["window_attention/y"])


def load_faf(faf, ref, pre="faf"):
    for name in ("temporal_attn0", "temporal_attn1", "downsample0", "downsample1", "upsample0", "upsample1", "fusion_tail"):
        load_conv(getattr(faf, name), ref, f"{pre}/{name}")
    load_conv(faf.feature_fusion[0], ref, f"{pre}/feature_fusion.0")
    faf.feature_fusion[1].weight.data = t(ref[f"{pre}/feature_fusion.1.weight"]).reshape(1)
    for i in range(5):
        for j in range(2):
            load_conv(faf.res_blocks[i][j].body[0], ref, f"{pre}/res_blocks.{i}.{j}.body.0")
            load_conv(faf.res_blocks[i][j].body[2], ref, f"{pre}/res_blocks.{i}.{j}.body.2")


def test_reference_layers_faf_block(ref):
    x = ref["faf/x"]  # [F, H, W, nf]
    Fr, H, W, nf = x.shape
    faf = O.FAFBlock(nf, Fr)
    load_faf(faf, ref)
    assert abs(faf.feature_fusion[1].weight.item() - 0.1) < 1e-7
    feat = t(x).permute(0, 3, 1, 2)[None]  # [1, F, C, H, W]
    g, gate = faf.guided(feat)
    close(g[0].permute(0, 2, 3, 1), ref["faf/guided"])
    assert gate.min().item() >= 0.5  # sigmoid(|.|): the gate never attenuates below one half
    fused, _ = faf.fuse(t(ref["faf/guided"]).permute(0, 3, 1, 2)[None])
    close(fused[0].permute(1, 2, 0), ref["faf/fused"])
    close(faf(feat)[0].permute(1, 2, 0), ref["faf/y"])


def test_reference_layers_shift_mask_partition_and_reverse(ref):
    """layers/fba_net.py:149-184 (mask), :199-203 (cyclic shift), :113-124 / :126-137 (partition / reverse), :228-231 (shift back),
    executed by the reference with the attention swapped for `x * token_gain`."""
    H = W = 8
    win, shift, d = 4, 2, 8
    mask = O.shift_attn_mask(H, W, win, shift)
    assert np.array_equal(mask.numpy(), ref["layer_shifted/mask"])
    norm1 = torch.nn.LayerNorm(d)
    norm1.weight.data, norm1.bias.data = t(ref["layer_shifted/norm1.weight"]), t(ref["layer_shifted/norm1.bias"])
    y = norm1(t(ref["layer_shifted/x"])).view(1, H, W, d)
    y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    yw = O.window_partition(y, win)
    close(yw, ref["layer_shifted/windows"])
    aw = yw * t(ref["layer_shifted/token_gain"])
    y = torch.roll(O.window_reverse(aw, win, 1, H, W), shifts=(shift, shift), dims=(1, 2))
    close(y.reshape(H * W, d), ref["layer_shifted/attn_path"])


def test_reference_layers_plain_layer_as_written(ref):
    """The unshifted heads = 1 `ffn` layer runs as written; its tail returns 2 * mlp(norm2(skip + attention path)) (A-4: the oracle
    keeps the Uformer residuals instead).  The oracle layer is run with its MLP swapped for a probe, which yields `skip + attention
    path`; the reference's tail is then rebuilt on top of it."""
    H = W = 8
    d, win = 8, 4
    pre = "layer_plain/"
    lay = O.LeWinLayer(d, (H, W), 1, win, 0, 4.0, "tanh")
    for n in ("norm1", "norm2"):
        getattr(lay, n).weight.data, getattr(lay, n).bias.data = t(ref[pre + n + ".weight"]), t(ref[pre + n + ".bias"])
    load_attention(lay.attn, ref, pre + "attn/", win)
    seen = {}

    class Probe(torch.nn.Module):
        def forward(self, u, H, W):
            seen["normed"] = u
            return torch.zeros_like(u)

    lay.mlp = Probe()
    lay(t(ref[pre + "x"])[None])
    u = seen["normed"][0]
    act = O.gelu_fn("tanh")  # A-13: jax.nn.gelu's default
    for i in range(3):
        u = F.linear(u, t(ref[pre + f"mlp.{i}.weight"]), t(ref[pre + f"mlp.{i}.bias"]))
        if i < 2:
            u = act(u)
    close(2 * u, ref[pre + "y"])


# ---------------------------------------------------------------------------------------------------------------------------
# the model file (models/fba_net.py): structure as the reference's own constructor built it, wiring as its own __call__ ran it
# (tests/golden/make_golden_model.py)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,kw", [("cfg2_rgb160", dict()), ("cfg3_raw80", dict(in_channels=4, img_size=80))])
def test_reference_model_structure(cfg, kw):
    import json

    with open(os.path.join(os.path.dirname(GOLD), "model_structure_reference.json")) as f:
        want = json.load(f)[cfg]
    m = O.OracleBaseModel(**kw)
    E, cin = m.embed_dim, m.in_channels
    extra_tail_conv = E * 4 * E * 9 + 4 * E  # A-17: the reference builds a x2 tail (scale_pow_two=1); x4 needs one more conv E -> 4E
    assert sum(p.numel() for p in m.parameters()) == want["total_params"] + extra_tail_conv
    names = {n for n, _ in m.named_children()}
    assert names == set(want["modules"]) - {"pos_drop"}  # Dropout(0.0), no parameters (A-8)
    for name, row in want["modules"].items():
        if name == "pos_drop":
            continue
        sub = getattr(m, name)
        have = sum(p.numel() for p in sub.parameters())
        assert have == row["params"] + (extra_tail_conv if name == "tail" else 0), name
        if "layers" in row:  # an FBANetBlock (blocks/fba_net.py) of FBANetLayers
            assert len(sub.blocks) == row["depth"] == len(row["layers"])
            for lay, r in zip(sub.blocks, row["layers"]):
                assert (lay.dim, list(lay.res), lay.attn.heads, lay.win, lay.shift) == (row["dim"], row["input_resolution"], r["heads"], r["window"], r["shift"]), name
                assert lay.dim // lay.attn.heads == r["dim_head"] and lay.mlp.linear1[0].out_features == r["mlp_hidden"]
                assert list(lay.attn.relative_position_bias_table.shape) == r["bias_table"]
    assert cin == (4 if "raw" in cfg else 3)


class _Mix(torch.nn.Module):
    """torch mirror of make_golden_model.py's `Stub`, in the layouts the oracle passes (tokens [B,T,C] or images [B,C,H,W])."""

    def __init__(self, z, name, kind):
        super().__init__()
        self.M, self.b, self.kind = t(z[f"stub/{name}/M"]).float(), t(z[f"stub/{name}/b"]).float(), kind

    def mix(self, x):
        return torch.tanh(x @ self.M + self.b)

    def forward(self, x, H=None, W=None):
        if self.kind == "tokens":
            return self.mix(x)
        if self.kind == "image":
            return self.mix(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        if self.kind == "tail":
            return self.mix(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2).repeat_interleave(4, 2).repeat_interleave(4, 3)
        B, T, C = x.shape
        img = x.view(B, H, W, C)
        if self.kind == "down":
            img = img.view(B, H // 2, 2, W // 2, 2, C).mean(dim=(2, 4))
        else:
            img = img.repeat_interleave(2, 1).repeat_interleave(2, 2)
        return self.mix(img.reshape(B, -1, C))


class _Fusion(torch.nn.Module):
    def __init__(self, z):
        super().__init__()
        self.inner = _Mix(z, "fusion", "image")

    def guided(self, feat):
        return feat, None

    def fuse(self, g):  # [B,F,C,H,W]: frames concatenated along channels, frame-major
        B, Fr, C, H, W = g.shape
        return self.inner(g.reshape(B, Fr * C, H, W)), None


def test_reference_model_wiring():
    z = np.load(os.path.join(os.path.dirname(GOLD), "model_wiring_reference.npz"))
    x = t(z["x"])  # [F, S, S, 3]
    Fr, S = x.shape[0], x.shape[1]
    m = O.OracleBaseModel(num_frames=Fr, img_size=S, in_channels=3, embed_dim=4, window_length=2)
    for name in ("head", "body", "input_proj", "output_proj", "output_proj_2", "output_proj_HG2_0", "output_proj_HG2_1"):
        setattr(m, name, _Mix(z, name, "image"))
    m.tail = _Mix(z, "tail", "tail")
    m.fusion = _Fusion(z)
    for hg in ("HG1", "HG2"):
        for name in (f"{hg}_encoderlayer_0", f"{hg}_encoderlayer_1", f"conv_{hg}", f"{hg}_decoderlayer_0", f"{hg}_decoderlayer_1"):
            setattr(m, name, _Mix(z, name, "tokens"))
        for i in (0, 1):
            setattr(m, f"{hg}_downsample_{i}", _Mix(z, f"{hg}_downsample_{i}", "down"))
            setattr(m, f"{hg}_upsample_{i}", _Mix(z, f"{hg}_upsample_{i}", "up"))
    st = m.forward_stages(x.permute(0, 3, 1, 2)[None])
    img = lambda v: v[0].permute(1, 2, 0)
    close(img(st["fusion"]), z["out/fusion"])
    close(st["input_proj"][0], z["out/input_proj"])
    for hg in ("HG1", "HG2"):
        for stage, mod in (("conv0", f"{hg}_encoderlayer_0"), ("pool0", f"{hg}_downsample_0"), ("conv1", f"{hg}_encoderlayer_1"),
                           ("pool1", f"{hg}_downsample_1"), ("conv2", f"conv_{hg}"), ("up0", f"{hg}_upsample_0"),
                           ("deconv0", f"{hg}_decoderlayer_0"), ("up1", f"{hg}_upsample_1"), ("deconv1", f"{hg}_decoderlayer_1")):
            close(st[f"{hg}.{stage}"][0], z[f"out/{mod}"])
    close(st["HG2.deconv0_in"][0], z["out/output_proj_HG2_0"])  # cat[up0, conv1, up0_2, conv1_2] (:305)
    close(st["HG2.deconv1_in"][0], z["out/output_proj_HG2_1"])  # cat[up1, conv0, up1_2, conv0_2] (:309)
    close(st["output_proj"][0], z["out/output_proj"])
    close(img(st["output_proj_2"]), z["out/output_proj_2"])
    close(img(st["tail"]), z["out/tail"])


def faf_gpu_case(ref, pre="faf_gpu"):
    """A GPU-sized FAF case of the fixture: (oracle FAFBlock carrying the reference's weights, input [1,F,C,H,W], reference gate, output)."""
    from tests_golden_helpers import det_array, faf_gpu_weights

    nf, frames, side = int(ref[pre + "/nf"]), int(ref[pre + "/frames"]), int(ref[pre + "/side"])
    w = {f"{pre}/{k}": v for k, v in faf_gpu_weights(nf, frames, int(ref[pre + "/seed0"])).items()}
    w[pre + "/feature_fusion.1.weight"] = np.float32(0.1)  # nn.PReLU(init_alpha=0.1), federated_affinity_fusion.py:47
    faf = O.FAFBlock(nf, frames)
    load_faf(faf, w, pre)
    x = det_array((frames, side, side, nf), int(ref[pre + "/x_seed"]), -1.0, 1.0)
    return faf, t(x).permute(0, 3, 1, 2)[None].contiguous(), ref[pre + "/gate"], ref[pre + "/y"]


@pytest.mark.parametrize("pre", ["faf_gpu", "faf_gpu64"])
def test_reference_layers_faf_block_gpu_size(ref, pre):
    faf, feat, gate, y = faf_gpu_case(ref, pre)
    with torch.no_grad():
        g, gt = faf.guided(feat)
        out, _ = faf.fuse(g)
    np.testing.assert_allclose(gt[0].numpy(), gate, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out[0].permute(1, 2, 0).numpy(), y, rtol=1e-4, atol=1e-4)


def attn_gpu_case(ref):
    """(oracle WindowAttention carrying the reference's weights, windows [nW, N, d], the reference's output [nW, N, d])."""
    from tests_golden_helpers import det_array

    dim, win, nwin, s0 = (int(ref["attn_gpu/" + k]) for k in ("dim", "win", "nwin", "seed0"))
    lim = 1.0 / np.sqrt(dim)
    w = {"a/qkv.to_q.weight": det_array((dim, dim), s0, -lim, lim), "a/qkv.to_q.bias": det_array((dim,), s0 + 1, -lim, lim),
         "a/qkv.to_kv.weight": det_array((2 * dim, dim), s0 + 2, -lim, lim), "a/qkv.to_kv.bias": det_array((2 * dim,), s0 + 3, -lim, lim),
         "a/proj.weight": det_array((dim, dim), s0 + 4, -lim, lim), "a/proj.bias": det_array((dim,), s0 + 5, -lim, lim),
         "a/relative_position_bias_table": det_array(((2 * win - 1) ** 2, 1), s0 + 6, -1.0, 1.0),
         "a/relative_position_index_as_written": ref["attn_gpu/index_as_written"]}
    attn = O.WindowAttention(dim, win, 1)
    load_attention(attn, w, "a/", win)
    return attn, t(det_array((nwin, win * win, dim), s0 + 7, -2.0, 2.0)), ref["attn_gpu/y"]


def qkv_gpu_case(ref):
    from tests_golden_helpers import det_array

    dim, heads, s0 = (int(ref["qkv_gpu/" + k]) for k in ("dim", "heads", "seed0"))
    lim = 1.0 / np.sqrt(dim)
    lp = O.LinearProjection(dim, heads)
    lp.load_state_dict({"to_q.weight": t(det_array((dim, dim), s0, -lim, lim)), "to_q.bias": t(det_array((dim,), s0 + 1, -lim, lim)),
                        "to_kv.weight": t(det_array((2 * dim, dim), s0 + 2, -lim, lim)), "to_kv.bias": t(det_array((2 * dim,), s0 + 3, -lim, lim))})
    return lp, t(det_array((100, dim), s0 + 4, -1.0, 1.0)), ref["qkv_gpu/q"], ref["qkv_gpu/k"], ref["qkv_gpu/v"]


def test_reference_layers_attention_and_qkv_gpu_size(ref):
    attn, xw, y = attn_gpu_case(ref)
    with torch.no_grad():
        np.testing.assert_allclose(attn(xw).numpy(), y, rtol=1e-4, atol=1e-4)
    lp, x, q, k, v = qkv_gpu_case(ref)
    with torch.no_grad():
        got = lp(x[None])
    for g, r in zip(got, (q, k, v)):
        np.testing.assert_allclose(g[0].numpy(), r, rtol=1e-4, atol=1e-5)


def test_reference_layers_pixel_shuffle_order(ref):
    """A-18: the reference's `"h w (h2 w2) -> (h h2) (w w2)"` on its one executable case (4 channels -> 1) is torch's PixelShuffle(2)
    order, in-channel `2 i + j` -> out `(2 y + i, 2 x + j)`, which the oracle's tail (and the CUDA tail's folded row permutation) use."""
    x = ref["pixel_shuffle/x"]  # [h, w, 4]
    y = torch.nn.PixelShuffle(2)(to_chw(x))  # [1, 1, 2h, 2w]
    assert np.array_equal(y[0, 0].numpy(), ref["pixel_shuffle/y"])
