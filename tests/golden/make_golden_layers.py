#!/usr/bin/env python
"""Golden vectors produced by EXECUTING the reference's own model code, layer by layer (build container only).

The reference's model is JAX/Equinox; neither library is installed and the whole-model forward cannot execute even with them
(SURVEY F2).  But most of its layers and blocks CAN run once `jax` / `equinox` resolve to something: `tests/golden/jaxshim/`
restates the documented semantics of the ~20 primitives the layers call in numpy, and this script imports the UNMODIFIED reference
packages `fba_net.layers.*` / `fba_net.blocks.*` from /root/reference on top of it, runs every layer that is executable as
written on seeded inputs and stores (weights as the reference holds them, inputs, outputs) in `layers_reference.npz`.
`tests/test_oracle_reference_layers.py` loads the weights into the oracle's modules (converting the layouts as SURVEY
A-11 says) and must reproduce the outputs.  What gets pinned is the reference's own composition code: axis swaps around the
convolutions, PReLU placement, the q / k / v split pattern, the attention arithmetic (heads = 1, the only case its asserts allow),
the FAF gate and hourglass wiring, window partition / cyclic shift / reverse and the 9-region mask construction.

Run from the repo root: ``python tests/golden/make_golden_layers.py``."""
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference")

import jax  # noqa: E402  (the shim)
from jax import numpy as jnp  # noqa: E402

from fba_net.blocks.federated_affinity_fusion import FAFBlock  # noqa: E402
from fba_net.blocks.residual import ResBlock  # noqa: E402
from fba_net.layers.conv2d import Conv2dLayer  # noqa: E402
from fba_net.layers.downsample_flatten import DownsampleFlattenLayer  # noqa: E402
from fba_net.layers.fba_net import FBANetLayer  # noqa: E402
from fba_net.layers.input_projection import InputProjLayer  # noqa: E402
from fba_net.layers.linear_projection import LinearProjectionLayer  # noqa: E402
from fba_net.layers.upsample_flatten import UpsampleFlattenLayer  # noqa: E402
from fba_net.layers.window_attention import WindowAttentionLayer  # noqa: E402

assert jax.__file__.startswith(HERE), "the shim must be the `jax` that got imported"
rng = np.random.default_rng(2024)
out = {}


def rand(*shape, lo=-1.0, hi=1.0):
    return jnp.asarray(rng.uniform(lo, hi, size=shape).astype(np.float32))


def put(prefix, **arrays):
    for k, v in arrays.items():
        out[f"{prefix}/{k}"] = np.asarray(v)


def conv_of(seq):
    """`Conv2dLayer` / `ConvTranspose2dLayer` return Sequential([swap, conv, swap]) (layers/conv2d.py:29-46)."""
    return seq[1]


def wb(prefix, name, conv):
    put(prefix, **{f"{name}.weight": conv.weight, f"{name}.bias": conv.bias})


def randomize_bias_like(layer):
    """Equinox initialises LayerNorm to (1, 0) and the bias table to N(0, 0.02): draw visible values so the test can see them."""
    layer.weight = rand(*layer.weight.shape, lo=0.5, hi=1.5)
    layer.bias = rand(*layer.bias.shape, lo=-0.5, hi=0.5)


# ---- convolutions (layers/conv2d.py, downsample_flatten.py, upsample_flatten.py): non-square maps, so the H/W swap shows
for name, layer, x in (
    ("conv3x3", Conv2dLayer(in_channels=3, out_channels=5), rand(6, 7, 3)),
    ("conv1x1", Conv2dLayer(in_channels=6, out_channels=4, kernel_size=1, padding=0), rand(5, 3, 6)),
    ("conv4x4s2", DownsampleFlattenLayer(in_channels=4, out_channels=6), rand(8, 6, 4)),
    ("depthwise", Conv2dLayer(in_channels=6, out_channels=6, groups=6, padding=(1, 1)), rand(5, 7, 6)),
    ("convT2x2s2", UpsampleFlattenLayer(in_channels=4, out_channels=3), rand(3, 5, 4)),
):
    wb(name, "conv", conv_of(layer))
    put(name, x=x, y=layer(x))

# ---- ResBlock (blocks/residual.py:21-29)
rb = ResBlock(num_feats=4)
x = rand(6, 5, 4)
wb("resblock", "body.0", conv_of(rb.body[0]))
wb("resblock", "body.2", conv_of(rb.body[2]))
put("resblock", x=x, y=rb(x))

# ---- InputProjLayer (layers/input_projection.py:30-46): conv3x3 + scalar PReLU(0.25) + flatten to tokens
ip = InputProjLayer(in_channels=4, out_channels=6)
x = rand(6, 6, 4)
wb("input_proj", "proj.0", conv_of(ip.body[0]))
put("input_proj", **{"proj.1.weight": ip.body[1].fn.negative_slope}, x=x, y=ip(x))

# ---- LinearProjectionLayer (layers/linear_projection.py:24-44): the "n (hd h c) -> hd h n c" split of to_kv
for heads in (1, 2, 4):
    lp = LinearProjectionLayer(dim=8, heads=heads)
    x = rand(9, 8)
    q, k, v = lp(x)
    put(f"linear_projection_h{heads}", **{"to_q.weight": lp.to_q.weight, "to_q.bias": lp.to_q.bias, "to_kv.weight": lp.to_kv.weight,
                                           "to_kv.bias": lp.to_kv.bias}, x=x, q=q, k=k, v=v)

# ---- WindowAttentionLayer (layers/window_attention.py:140-248), heads = 1 (its asserts at :175-177 hold for no other value), no mask
wa = WindowAttentionLayer(dim=8, window_length=3, heads=1)
wa.relative_position_bias_table = rand(*wa.relative_position_bias_table.shape, lo=-0.5, hi=0.5)
x = rand(9, 8)


def attn_weights(prefix, a):
    put(prefix, **{"relative_position_bias_table": a.relative_position_bias_table, "relative_position_index_as_written": a.relative_position_index,
                   "qkv.to_q.weight": a.qkv.to_q.weight, "qkv.to_q.bias": a.qkv.to_q.bias, "qkv.to_kv.weight": a.qkv.to_kv.weight,
                   "qkv.to_kv.bias": a.qkv.to_kv.bias, "proj.weight": a.proj.weight, "proj.bias": a.proj.bias})


attn_weights("window_attention", wa)
put("window_attention", x=x, y=wa(x))

# ---- FAFBlock (blocks/federated_affinity_fusion.py:18-182), whole block, non-square map
nf, frames = 4, 3
faf = FAFBlock(num_feats=nf, num_frames=frames)
x = rand(frames, 8, 12, nf)
for name in ("temporal_attn0", "temporal_attn1", "downsample0", "downsample1", "upsample0", "upsample1", "fusion_tail"):
    wb("faf", name, conv_of(getattr(faf, name)))
wb("faf", "feature_fusion.0", conv_of(faf.feature_fusion[0]))
put("faf", **{"feature_fusion.1.weight": faf.feature_fusion[1].fn.negative_slope})
for i, seq in enumerate(faf.res_blocks):
    for j in range(2):
        blk = seq[j].fn
        wb("faf", f"res_blocks.{i}.{j}.body.0", conv_of(blk.body[0]))
        wb("faf", f"res_blocks.{i}.{j}.body.2", conv_of(blk.body[2]))
guided = faf.compute_guided_aligned_features(x)
put("faf", x=x, guided=guided, fused=faf.fuse_features(guided), y=faf(x))

# ---- FBANetLayer (layers/fba_net.py:19-250)
# (1) shifted layer, attention and drop_path swapped for recorders: the reference's own mask construction (:149-184), cyclic shift,
#     window partition, window reverse and shift back run as written; the `assert False` of window_attention.py:215 is never reached.
H, W, d, win, shift = 8, 8, 8, 4, 2
lay = FBANetLayer(dim=d, input_resolution=(H, W), heads=1, window_height=win, window_width=win, shift_size_height=shift,
                  shift_size_width=shift, token_mlp="ffn")
randomize_bias_like(lay.norm1)
seen = {}
token_gain = (1.0 + 0.125 * np.arange(win * win, dtype=np.float32))[:, None]


def attn_recorder(xw, mask=None):
    seen.setdefault("windows", []).append(np.asarray(xw))
    seen["mask"] = np.asarray(mask)
    return xw * token_gain                     # depends on the position inside the window: pins reverse + shift back


def drop_path_recorder(v):
    seen.setdefault("drop_path_in", []).append(np.asarray(v))
    return v


lay.attn = attn_recorder
lay.drop_path = drop_path_recorder
x = rand(H * W, d)
lay(x)
put("layer_shifted", **{"norm1.weight": lay.norm1.weight, "norm1.bias": lay.norm1.bias}, x=x, mask=seen["mask"],
    windows=np.stack(seen["windows"]), token_gain=token_gain, attn_path=seen["drop_path_in"][0])

# (2) unshifted layer, heads = 1, token_mlp = "ffn": executable exactly as written.  Its tail (:244-248) computes
#     2 * mlp(norm2(skip + attn_path)) -- SURVEY A-4; the test rebuilds that from the oracle's `skip + attn_path`.
lay = FBANetLayer(dim=d, input_resolution=(H, W), heads=1, window_height=win, window_width=win, token_mlp="ffn")
randomize_bias_like(lay.norm1)
randomize_bias_like(lay.norm2)
lay.attn.relative_position_bias_table = rand(*lay.attn.relative_position_bias_table.shape, lo=-0.5, hi=0.5)
x = rand(H * W, d)
attn_weights("layer_plain/attn", lay.attn)
put("layer_plain", **{"norm1.weight": lay.norm1.weight, "norm1.bias": lay.norm1.bias, "norm2.weight": lay.norm2.weight,
                      "norm2.bias": lay.norm2.bias, "mlp.0.weight": lay.mlp.layers[0].weight, "mlp.0.bias": lay.mlp.layers[0].bias,
                      "mlp.1.weight": lay.mlp.layers[1].weight, "mlp.1.bias": lay.mlp.layers[1].bias,
                      "mlp.2.weight": lay.mlp.layers[2].weight, "mlp.2.bias": lay.mlp.layers[2].bias}, x=x, y=lay(x))

# ---- FAFBlock once more at a size the CUDA path runs (num_feats 32, 4 frames, 40 x 40 = tests/test_gpu_model.py SMALL): weights and input
# come from `det_array` (rebuilt by the test from the seeds), only the reference's output is stored.  `-m gpu` compares the CUDA FAF
# (gate kernel + implicit GEMMs) with it directly, not via the oracle.
sys.path.insert(0, os.path.dirname(HERE))
from tests_golden_helpers import det_array, faf_gpu_weights  # noqa: E402

def faf_gpu_case(prefix, nf, frames, side, seed0, x_seed):
    faf = FAFBlock(num_feats=nf, num_frames=frames)
    weights = faf_gpu_weights(nf, frames, seed0)

    def set_conv(conv, name):
        assert conv.weight.shape == weights[name + ".weight"].shape and conv.bias.shape == weights[name + ".bias"].shape, name
        conv.weight, conv.bias = jnp.asarray(weights.pop(name + ".weight")), jnp.asarray(weights.pop(name + ".bias"))

    for name in ("temporal_attn0", "temporal_attn1", "downsample0", "downsample1", "upsample0", "upsample1", "fusion_tail"):
        set_conv(conv_of(getattr(faf, name)), name)
    set_conv(conv_of(faf.feature_fusion[0]), "feature_fusion.0")
    for i, seq in enumerate(faf.res_blocks):
        for j in range(2):
            for k in (0, 2):
                set_conv(conv_of(seq[j].fn.body[k]), f"res_blocks.{i}.{j}.body.{k}")
    assert not weights
    x = jnp.asarray(det_array((frames, side, side, nf), x_seed, -1.0, 1.0))
    guided = faf.compute_guided_aligned_features(x)
    g_, x_ = np.asarray(guided, dtype=np.float64)[1:], np.asarray(x, dtype=np.float64)[1:]
    put(prefix, gate=(g_ * x_).sum(-1) / (x_ * x_).sum(-1), y=faf(x), nf=nf, frames=frames, side=side, seed0=seed0, x_seed=x_seed)


faf_gpu_case("faf_gpu", 32, 4, 40, 1000, 999)      # fp32 CUDA-core path (tests/test_gpu_model.py SMALL)
faf_gpu_case("faf_gpu64", 64, 4, 40, 2000, 1999)   # 64-channel granularity: the tcgen05 path in bf16

# ---- WindowAttentionLayer at the size of the model's first encoder stage (dim 64, window 10, heads 1 -- models/fba_net.py:130-139 with
# get_arch's window), four windows through `jax.vmap` as layers/fba_net.py:222 does; weights / windows from `det_array`.
# And LinearProjectionLayer at the second stage's size (dim 128, heads 2): the q / k / v column layout the CUDA qkv GEMM must produce.
dim, win, nwin = 64, 10, 4
wa = WindowAttentionLayer(dim=dim, window_length=win, heads=1)
lim = 1.0 / np.sqrt(dim)
wa.qkv.to_q.weight, wa.qkv.to_q.bias = jnp.asarray(det_array((dim, dim), 3000, -lim, lim)), jnp.asarray(det_array((dim,), 3001, -lim, lim))
wa.qkv.to_kv.weight, wa.qkv.to_kv.bias = jnp.asarray(det_array((2 * dim, dim), 3002, -lim, lim)), jnp.asarray(det_array((2 * dim,), 3003, -lim, lim))
wa.proj.weight, wa.proj.bias = jnp.asarray(det_array((dim, dim), 3004, -lim, lim)), jnp.asarray(det_array((dim,), 3005, -lim, lim))
wa.relative_position_bias_table = jnp.asarray(det_array(((2 * win - 1) ** 2, 1), 3006, -1.0, 1.0))
xw = jnp.asarray(det_array((nwin, win * win, dim), 3007, -2.0, 2.0))
put("attn_gpu", y=jax.vmap(wa)(xw), index_as_written=wa.relative_position_index, dim=dim, win=win, nwin=nwin, seed0=3000)

dim2, heads2 = 128, 2
lp = LinearProjectionLayer(dim=dim2, heads=heads2)
lim = 1.0 / np.sqrt(dim2)
lp.to_q.weight, lp.to_q.bias = jnp.asarray(det_array((dim2, dim2), 3100, -lim, lim)), jnp.asarray(det_array((dim2,), 3101, -lim, lim))
lp.to_kv.weight, lp.to_kv.bias = jnp.asarray(det_array((2 * dim2, dim2), 3102, -lim, lim)), jnp.asarray(det_array((2 * dim2,), 3103, -lim, lim))
q, k, v = lp(jnp.asarray(det_array((100, dim2), 3104, -1.0, 1.0)))
put("qkv_gpu", q=q, k=k, v=v, dim=dim2, heads=heads2, seed0=3100)

# ---- PixelShuffleLayer (layers/pixel_shuffle.py:9-10): its pattern drops the channel axis, so it only runs on exactly 4 channels
# (SURVEY A-18) -- enough to pin which of the 4 sub-pixels lands where
from fba_net.layers.pixel_shuffle import PixelShuffleLayer  # noqa: E402

x = rand(3, 5, 4)
put("pixel_shuffle", x=x, y=PixelShuffleLayer(2)(x))

np.savez_compressed(os.path.join(HERE, "layers_reference.npz"), **{k: np.asarray(v, dtype=np.float32) if np.asarray(v).dtype.kind == "f" else np.asarray(v)
                                                                  for k, v in out.items()})
print("reference layer fixtures:", len(out), "arrays,", sum(np.asarray(v).size for v in out.values()), "values")
