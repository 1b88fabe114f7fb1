#!/usr/bin/env python
"""Golden data produced by EXECUTING the reference's model file `fba_net/models/fba_net.py` (build container only), on top of the
numpy stand-ins of `tests/golden/jaxshim/` (see make_golden_layers.py for the layers themselves).

1. **Structure** (`model_structure_reference.json`): `FBANetModel(...)` is constructed exactly as `utils/model_utils.py:65-82` `get_arch`
   does for the BASELINE configuration (img_size 160, embed_dim 64, window 10, linear projection, LeFF) and for the RAW one (80, 4
   input channels); the tree the reference's own `__post_init__` built (models/fba_net.py:81-235) is walked and every block's
   dim / resolution / depth / heads / window / shift / drop-path rate and every attribute's parameter count are written down.
2. **Wiring** (`model_wiring_reference.npz`): the reference's own `__call__` (models/fba_net.py:242-322) is executed on a small model
   whose sub-modules were swapped for cheap deterministic stubs (a channel-mixing matrix + tanh, 2x2 mean pooling for the
   downsamplers, pixel repetition for the upsamplers).  What runs is the reference's wiring: which module consumes what, the
   concatenation orders of :282/:286/:305/:309, HG2's reuse of HG1's tensors.  The call stops after the tail at :317 (`jim.resize`
   with a x4 channel axis, SURVEY A-19) -- everything up to `output_2` is recorded.

Run from the repo root: ``python tests/golden/make_golden_model.py``."""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference")

from jax import numpy as jnp  # noqa: E402  (the shim)

from fba_net.blocks.fba_net import FBANetBlock  # noqa: E402
from fba_net.models.fba_net import FBANetModel  # noqa: E402


# ------------------------------------------------------------------------------------------------ 1. structure
def n_params(o, seen):
    if id(o) in seen:
        return 0
    seen.add(id(o))
    if isinstance(o, np.ndarray):
        return int(o.size) if o.dtype.kind == "f" else 0
    if isinstance(o, (list, tuple)):
        return sum(n_params(v, seen) for v in o)
    if hasattr(o, "__dict__"):
        return sum(n_params(v, seen) for v in vars(o).values())
    return 0


def describe(model):
    rows = {}
    for name, sub in vars(model).items():
        if not hasattr(sub, "__dict__"):
            continue
        row = {"params": n_params(sub, set())}
        if isinstance(sub, FBANetBlock):
            layers = [lam.fn for lam in sub.body.layers]
            row.update(dim=sub.dim, input_resolution=list(sub.input_resolution), depth=sub.depth, heads=sub.heads,
                       layers=[dict(window=lay.window_height, shift=lay.shift_size_height, heads=lay.heads, dim_head=lay.attn.dim_head,
                                    bias_table=list(lay.attn.relative_position_bias_table.shape),
                                    mlp_hidden=int(lay.mlp.hidden_dim), drop_path_rate=round(float(lay.drop_path_rate), 6))
                               for lay in layers])
        rows[name] = row
    return {"total_params": n_params(model, set()), "modules": rows}


structure = {
    "cfg2_rgb160": describe(FBANetModel(img_size=160, embed_dim=64, window_length=10, token_projection="linear", token_mlp="leff")),
    "cfg3_raw80": describe(FBANetModel(img_size=80, in_channels=4, embed_dim=64, window_length=10, token_projection="linear", token_mlp="leff")),
}
with open(os.path.join(HERE, "model_structure_reference.json"), "w") as f:
    json.dump(structure, f, indent=1, sort_keys=True)
print("reference structure:", {k: v["total_params"] for k, v in structure.items()})

# ------------------------------------------------------------------------------------------------ 2. wiring
rng = np.random.default_rng(77)
F_, S, E = 3, 8, 4
out = {}
order = []


class Stub:
    """y = tanh(pre(x) @ M + b) over the last axis.  `kind`: "mix" (shape kept), "frames" ([F,H,W,C] -> [H,W,F*C] first),
    "flatten" ([H,W,C] -> [H*W,C] afterwards), "down" (tokens, 2x2 mean pool first), "up" (tokens, 2x2 pixel repetition first),
    "image" (tokens -> [S,S,C] afterwards)."""

    def __init__(self, name, kind, cin, cout):
        self.name, self.kind = name, kind
        self.M = rng.uniform(-1, 1, size=(cin, cout)).astype(np.float32) / np.sqrt(cin)
        self.b = rng.uniform(-0.2, 0.2, size=(cout,)).astype(np.float32)
        out[f"stub/{name}/M"], out[f"stub/{name}/b"] = self.M, self.b

    def __call__(self, x, **kw):
        x = np.asarray(x)
        if self.kind == "frames":
            x = np.concatenate(list(x), axis=-1)
        if self.kind in ("down", "up"):
            side = int(round(np.sqrt(x.shape[0])))
            x = x.reshape(side, side, -1)
            if self.kind == "down":
                x = x.reshape(side // 2, 2, side // 2, 2, -1).mean(axis=(1, 3))
            else:
                x = x.repeat(2, axis=0).repeat(2, axis=1)
            x = x.reshape(-1, x.shape[-1])
        y = np.tanh(x @ self.M + self.b)
        if self.kind == "flatten":
            y = y.reshape(-1, y.shape[-1])
        if self.kind == "image":
            side = int(round(np.sqrt(y.shape[0])))
            y = y.reshape(side, side, -1)
        if self.kind == "tail":
            y = y.repeat(4, axis=0).repeat(4, axis=1)
        out[f"out/{self.name}"] = y.astype(np.float32)
        if self.name not in order:
            order.append(self.name)
        return jnp.asarray(y)


m = FBANetModel(num_frames=F_, img_size=S, embed_dim=E, window_length=2, token_projection="linear", token_mlp="leff")
stubs = dict(head=("mix", 3, E), body=("mix", E, E), fusion=("frames", F_ * E, E), input_proj=("flatten", E, E),
             output_proj=("mix", 2 * E, E), output_proj_2=("image", 2 * E, E), output_proj_HG2_0=("mix", 8 * E, 4 * E),
             output_proj_HG2_1=("mix", 4 * E, 2 * E), tail=("tail", E, 3))
for hg in ("HG1", "HG2"):
    stubs.update({f"{hg}_encoderlayer_0": ("mix", E, E), f"{hg}_downsample_0": ("down", E, 2 * E), f"{hg}_encoderlayer_1": ("mix", 2 * E, 2 * E),
                  f"{hg}_downsample_1": ("down", 2 * E, 4 * E), f"conv_{hg}": ("mix", 4 * E, 4 * E), f"{hg}_upsample_0": ("up", 4 * E, 2 * E),
                  f"{hg}_decoderlayer_0": ("mix", 4 * E, 4 * E), f"{hg}_upsample_1": ("up", 4 * E, E), f"{hg}_decoderlayer_1": ("mix", 2 * E, 2 * E)})
for name, (kind, cin, cout) in stubs.items():
    assert hasattr(m, name), name
    setattr(m, name, Stub(name, kind, cin, cout))
x = jnp.asarray(rng.uniform(0, 1, size=(F_, S, S, 3)).astype(np.float32))
out["x"] = np.asarray(x)
try:
    m(x)
    raise SystemExit("the reference forward was expected to stop at models/fba_net.py:317")
except NotImplementedError:
    pass
assert order[-1] == "tail" and len(order) == len(stubs), order
out["call_order"] = np.array(order)
np.savez_compressed(os.path.join(HERE, "model_wiring_reference.npz"), **out)
print("reference wiring: modules called in order:", " ".join(order))
