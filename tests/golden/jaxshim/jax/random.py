"""`jax.random` subset: keys are numpy SeedSequences (values differ from JAX's threefry; only determinism matters here)."""
import numpy as _np


def key(seed):
    return _np.random.SeedSequence(seed)


PRNGKey = key


def split(k, num=2):
    return tuple(k.spawn(num))


def _rng(k):
    return _np.random.default_rng(k)


def uniform(key, shape=(), dtype=_np.float32, minval=0.0, maxval=1.0):
    return _rng(key).uniform(minval, maxval, size=shape).astype(dtype)


def truncated_normal(key, lower, upper, shape=(), dtype=_np.float32):
    r = _rng(key)
    out = r.standard_normal(size=shape)
    bad = (out < lower) | (out > upper)
    while bad.any():
        out[bad] = r.standard_normal(size=int(bad.sum()))
        bad = (out < lower) | (out > upper)
    return out.astype(dtype)


def bernoulli(key, p=0.5, shape=()):
    return _rng(key).uniform(size=shape) < p
