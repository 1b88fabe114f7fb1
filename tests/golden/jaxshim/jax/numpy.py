"""`jax.numpy` subset on numpy: ndarray subclass with `.at[idx].set(v)` (functional update) and jnp's clamped gather."""
import numpy as _np
from numpy import abs, arange, concatenate, expand_dims, float32, int32, mgrid, ones, roll as _roll, stack, sum, where  # noqa: F401,A004


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            @staticmethod
            def set(v):
                out = _np.array(arr, copy=True)
                out[idx] = v
                return asarray(out)

        return _Setter()


class Array(_np.ndarray):
    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        # jnp gathers clamp out-of-bounds integer-array indices (NumPy would raise): jax docs, "out-of-bounds indexing"
        if isinstance(idx, _np.ndarray) and idx.dtype.kind in "iu":
            idx = _np.clip(_np.asarray(idx), -self.shape[0], self.shape[0] - 1)
        return super().__getitem__(idx)


def asarray(x, dtype=None):
    return _np.asarray(x, dtype=dtype).view(Array)


array = asarray


def zeros(shape, dtype=_np.float32):
    return asarray(_np.zeros(shape, dtype=dtype))


def roll(a, shift, axis=None):
    return asarray(_roll(a, shift, axis))


def linspace(start, stop, num=50):
    return asarray(_np.linspace(start, stop, num, dtype=_np.float32))
