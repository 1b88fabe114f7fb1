#!/usr/bin/env python
"""Coefficients of the FMA-pipe GELU used by fbanet_b200/csrc/leff_mlp_tcgen05.cu (`gelu_half_poly_f2`).

GELU_tanh(x) = relu(x) - q(|x|) with q(a) = a / (1 + exp(2 k0 (a + k1 a^3))): a smooth bump, < 7e-5 beyond a = 4.  q is fitted on
[0, 4] by a degree-8 polynomial in t = a/2 - 1 (Chebyshev nodes, converted to the monomial basis for Horner evaluation); the script
prints the coefficients (highest first) and the max abs error of the fp32 Horner evaluation of the whole GELU over [-8, 8]."""
import numpy as np
from numpy.polynomial import chebyshev as C

K0, K1 = 0.7978845608028654, 0.044715


def q(a):
    return a / (1.0 + np.exp(2.0 * K0 * (a + K1 * a ** 3)))


def gelu(x):
    return 0.5 * x * (1.0 + np.tanh(K0 * (x + K1 * x ** 3)))


deg, n = 8, 600
k = np.cos(np.pi * (np.arange(n) + 0.5) / n)
cheb = C.chebfit(k, q(2.0 * (k + 1.0)), deg)
mono = C.cheb2poly(cheb)                      # lowest first
print("constexpr float " + ", ".join(f"c{i} = {mono[i]:.8e}f" for i in range(deg, -1, -1)) + ";")
x = np.linspace(-8, 8, 400001).astype(np.float32)
z = (x * np.float32(0.5)).astype(np.float32)
t = (np.minimum(np.abs(z), np.float32(2.0)) - np.float32(1.0)).astype(np.float32)
r = np.full_like(t, np.float32(mono[deg]))
for i in range(deg - 1, -1, -1):
    r = (r * t + np.float32(mono[i])).astype(np.float32)
y = np.maximum(z + z, np.float32(0)) - r
err = np.abs(y.astype(np.float64) - gelu(x.astype(np.float64)))
print(f"max abs error of the fp32 evaluation on [-8, 8]: {err.max():.3e} at x = {x[err.argmax()]:.3f}")
