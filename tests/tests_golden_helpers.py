"""Shared loader of the fixtures produced by executing the reference's own code (tests/golden/make_golden_reference.py)."""
import os

import numpy as np


def tiling_reference():
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiling_reference.npz"))
    n, C, side = d["tiles"].shape[0], d["tiles"].shape[2], int(d["scale"]) * (int(d["psize"]) + 2 * int(d["overlap"]))
    k, c, y, x = np.meshgrid(np.arange(n), np.arange(C), np.arange(side), np.arange(side), indexing="ij")
    sr = (((k * 7 + c * 3 + y * 5 + x * 11) % 97) / 97.0).astype(np.float32)       # same pattern as make_golden_reference.py
    return d, sr


def det_array(shape, seed, lo=-1.0, hi=1.0):
    """Deterministic pseudo-random float32 array (splitmix64 over the flat index): bit-identical wherever numpy runs, so large
    weights / inputs of a reference-executed fixture need not be stored -- generator and test both rebuild them from (shape, seed)."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


# FAF block of the GPU-sized fixture (tests/golden/make_golden_layers.py, `faf_gpu/*`): every conv of the reference's FAFBlock in the
# order its fields are declared, with the seeds its weights / biases were drawn from.  name -> (weight shape AS THE REFERENCE HOLDS IT)
def faf_gpu_spec(nf, frames):
    spec = [("temporal_attn0", (nf, nf, 3, 3)), ("temporal_attn1", (nf, nf, 3, 3)), ("feature_fusion.0", (nf, nf * frames, 1, 1)),
            ("downsample0", (2 * nf, nf, 4, 4)), ("downsample1", (4 * nf, 2 * nf, 4, 4)),
            ("upsample0", (2 * nf, 4 * nf, 2, 2)), ("upsample1", (nf, 4 * nf, 2, 2))]  # Equinox ConvTranspose2d: [out, in, kh, kw]
    for i, mult in enumerate((1, 2, 4, 4, 2)):
        for j in range(2):
            for k in (0, 2):
                spec.append((f"res_blocks.{i}.{j}.body.{k}", (nf * mult, nf * mult, 3, 3)))
    spec.append(("fusion_tail", (nf, 2 * nf, 3, 3)))
    return spec


def faf_gpu_weights(nf, frames, seed0=1000):
    """{name.weight / name.bias: array} in the REFERENCE's layouts, fan-in scaled like its initialisers."""
    out = {}
    for n, (name, shape) in enumerate(faf_gpu_spec(nf, frames)):
        lim = 1.0 / np.sqrt(shape[1] * shape[2] * shape[3])
        out[name + ".weight"] = det_array(shape, seed0 + 2 * n, -lim, lim)
        out[name + ".bias"] = det_array((shape[0], 1, 1), seed0 + 2 * n + 1, -lim, lim)
    return out
