"""`jax.nn` subset.  `gelu` defaults to approximate=True (the tanh form), as in JAX."""
import numpy as _np


def relu(x):
    return _np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + _np.exp(-x))


def gelu(x, approximate=True):
    if approximate:
        return 0.5 * x * (1.0 + _np.tanh(_np.sqrt(2.0 / _np.pi) * (x + 0.044715 * x**3)))
    from scipy.special import erf

    return 0.5 * x * (1.0 + erf(x / _np.sqrt(2.0)))


def softmax(x, axis=-1):
    e = _np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)
