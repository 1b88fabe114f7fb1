"""Shared loader of the fixtures produced by executing the reference's own code (tests/golden/make_golden_reference.py)."""
import os

import numpy as np


def tiling_reference():
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiling_reference.npz"))
    n, C, side = d["tiles"].shape[0], d["tiles"].shape[2], int(d["scale"]) * (int(d["psize"]) + 2 * int(d["overlap"]))
    k, c, y, x = np.meshgrid(np.arange(n), np.arange(C), np.arange(side), np.arange(side), indexing="ij")
    sr = (((k * 7 + c * 3 + y * 5 + x * 11) % 97) / 97.0).astype(np.float32)       # same pattern as make_golden_reference.py
    return d, sr
