"""`equinox.nn` subset in numpy, each layer following Equinox's documented semantics (v0.11):

* `Linear(in, out, use_bias, key)`: `weight [out, in]`, `bias [out]`, both U(-1/sqrt(in), 1/sqrt(in)); call on ONE vector: `W @ x + b`.
* `Conv2d(...)`: cross-correlation on ONE `[C, H, W]` image; `weight [out, in/groups, kh, kw]`, `bias [out, 1, 1]`, zero padding.
* `ConvTranspose2d(...)`: `weight [out, in/groups, kh, kw]`; the input is dilated by the stride, padded by `dilation*(k-1) - p` and
  cross-correlated with the weight AS STORED (Equinox does not flip the kernel; torch's ConvTranspose2d weight is this one flipped
  along both spatial axes with the first two axes swapped).
* `LayerNorm(shape, eps=1e-5)`: biased variance over the whole input, `weight` ones, `bias` zeros.
* `PReLU(init_alpha=0.25)`: `where(x >= 0, x, negative_slope * x)`, scalar slope.
* `Sequential(layers)`, `Lambda(fn)`, `Identity()`, `Dropout(p)` (identity at p == 0 or inference), `MLP`."""
import math

import numpy as np
from jax import numpy as jnp
from jax import random as jrandom


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class _Layer:
    def __call__(self, x, *, key=None):
        raise NotImplementedError


class Identity(_Layer):
    def __init__(self, *a, **k):
        pass

    def __call__(self, x, *, key=None):
        return x


class Lambda(_Layer):
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, x, *, key=None):
        return self.fn(x)


class Sequential(_Layer):
    def __init__(self, layers):
        self.layers = tuple(layers)

    def __call__(self, x, *, state=None, key=None):
        for layer in self.layers:
            x = layer(x)
        return x

    def __getitem__(self, i):
        return self.layers[i]

    def __len__(self):
        return len(self.layers)


class Dropout(_Layer):
    def __init__(self, p=0.5, inference=False):
        self.p, self.inference = p, inference

    def __call__(self, x, *, key=None, inference=None):
        if self.inference or self.p == 0:
            return x
        raise RuntimeError("Dropout requires a key when running in non-deterministic mode.")


class PReLU(_Layer):
    def __init__(self, init_alpha=0.25):
        self.negative_slope = jnp.asarray(init_alpha, dtype=np.float32)

    def __call__(self, x, *, key=None):
        return np.where(x >= 0, x, self.negative_slope * x)


class Linear(_Layer):
    def __init__(self, in_features, out_features, use_bias=True, *, key):
        wkey, bkey = jrandom.split(key, 2)
        lim = 1 / math.sqrt(in_features)
        self.weight = jnp.asarray(jrandom.uniform(wkey, (out_features, in_features), minval=-lim, maxval=lim))
        self.bias = jnp.asarray(jrandom.uniform(bkey, (out_features,), minval=-lim, maxval=lim)) if use_bias else None
        self.in_features, self.out_features, self.use_bias = in_features, out_features, use_bias

    def __call__(self, x, *, key=None):
        assert x.shape == (self.in_features,), x.shape
        y = self.weight @ x
        return y + self.bias if self.bias is not None else y


class LayerNorm(_Layer):
    def __init__(self, shape, eps=1e-5, use_weight=True, use_bias=True):
        self.shape = (shape,) if isinstance(shape, int) else tuple(shape)
        self.eps = eps
        self.weight = jnp.asarray(np.ones(self.shape, np.float32)) if use_weight else None
        self.bias = jnp.zeros(self.shape) if use_bias else None

    def __call__(self, x, state=None, *, key=None):
        assert x.shape == self.shape, (x.shape, self.shape)
        mean = x.mean(keepdims=True)
        var = x.var(keepdims=True)  # biased
        out = (x - mean) / np.sqrt(np.maximum(var, 0.0) + self.eps)
        if self.weight is not None:
            out = self.weight * out
        if self.bias is not None:
            out = out + self.bias
        return out


def _correlate(x, w, stride, groups):
    """x [C, H, W] already padded / dilated; w [O, C/groups, kh, kw]; plain cross-correlation, tap by tap."""
    O, Cg, kh, kw = w.shape
    C, H, W = x.shape
    assert C == Cg * groups and O % groups == 0
    sh, sw = stride
    Ho, Wo = (H - kh) // sh + 1, (W - kw) // sw + 1
    out = np.zeros((O, Ho, Wo), dtype=np.result_type(x, w))
    Og = O // groups
    for g in range(groups):
        xg = x[g * Cg:(g + 1) * Cg]
        wg = w[g * Og:(g + 1) * Og]
        for ky in range(kh):
            for kx in range(kw):
                patch = xg[:, ky:ky + sh * (Ho - 1) + 1:sh, kx:kx + sw * (Wo - 1) + 1:sw]
                out[g * Og:(g + 1) * Og] += np.einsum("oc,chw->ohw", wg[:, :, ky, kx], patch)
    return out


class Conv2d(_Layer):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, use_bias=True, *, key):
        self.kernel_size, self.stride, self.dilation = _pair(kernel_size), _pair(stride), _pair(dilation)
        self.padding = _pair(padding)
        assert self.dilation == (1, 1)
        self.groups, self.in_channels, self.out_channels = groups, in_channels, out_channels
        wkey, bkey = jrandom.split(key, 2)
        cg = in_channels // groups
        lim = 1 / math.sqrt(cg * self.kernel_size[0] * self.kernel_size[1])
        self.weight = jnp.asarray(jrandom.uniform(wkey, (out_channels, cg) + self.kernel_size, minval=-lim, maxval=lim))
        self.bias = jnp.asarray(jrandom.uniform(bkey, (out_channels, 1, 1), minval=-lim, maxval=lim)) if use_bias else None

    def __call__(self, x, *, key=None):
        assert x.ndim == 3 and x.shape[0] == self.in_channels, x.shape
        ph, pw = self.padding
        xp = np.pad(np.asarray(x), ((0, 0), (ph, ph), (pw, pw)))
        y = _correlate(xp, np.asarray(self.weight), self.stride, self.groups)
        return jnp.asarray(y + self.bias if self.bias is not None else y)


class ConvTranspose2d(_Layer):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, dilation=1, groups=1,
                 use_bias=True, *, key):
        self.kernel_size, self.stride, self.padding = _pair(kernel_size), _pair(stride), _pair(padding)
        assert _pair(dilation) == (1, 1) and _pair(output_padding) == (0, 0) and groups == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        wkey, bkey = jrandom.split(key, 2)
        lim = 1 / math.sqrt(in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight = jnp.asarray(jrandom.uniform(wkey, (out_channels, in_channels) + self.kernel_size, minval=-lim, maxval=lim))
        self.bias = jnp.asarray(jrandom.uniform(bkey, (out_channels, 1, 1), minval=-lim, maxval=lim)) if use_bias else None

    def __call__(self, x, *, key=None):
        assert x.ndim == 3 and x.shape[0] == self.in_channels, x.shape
        C, H, W = x.shape
        sh, sw = self.stride
        z = np.zeros((C, (H - 1) * sh + 1, (W - 1) * sw + 1), dtype=x.dtype)  # lhs_dilation = stride
        z[:, ::sh, ::sw] = x
        ph, pw = (k - 1 - p for k, p in zip(self.kernel_size, self.padding))
        zp = np.pad(z, ((0, 0), (ph, ph), (pw, pw)))
        y = _correlate(zp, np.asarray(self.weight), (1, 1), 1)
        return jnp.asarray(y + self.bias if self.bias is not None else y)


class MLP(_Layer):
    def __init__(self, in_size, out_size, width_size, depth, activation=None, final_activation=None, use_bias=True,
                 use_final_bias=True, *, key):
        keys = jrandom.split(key, depth + 1)
        sizes = [in_size] + [width_size] * depth + [out_size]
        self.layers = tuple(Linear(sizes[i], sizes[i + 1], key=keys[i]) for i in range(depth + 1))
        self.activation = activation
        self.final_activation = final_activation or (lambda v: v)

    def __call__(self, x, *, key=None):
        for layer in self.layers[:-1]:
            x = self.activation(layer(x))  # Equinox vmaps the activation over the units; these are elementwise
        return self.final_activation(self.layers[-1](x))
