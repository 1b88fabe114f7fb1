"""numpy stand-in for the few `jax` names the reference's layers use (see ../README.md).  Build container only."""
import numpy as _np

from . import image, nn, numpy, random  # noqa: F401


def vmap(fn, in_axes=0, out_axes=0):
    """`jax.vmap` over the leading axis of every positional argument, outputs stacked (tuples stacked leaf by leaf)."""
    assert in_axes == 0 and out_axes == 0

    def mapped(*args):
        n = args[0].shape[0]
        outs = [fn(*[a[i] for a in args]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(numpy.asarray(_np.stack([o[j] for o in outs])) for j in range(len(outs[0])))
        return numpy.asarray(_np.stack(outs))

    return mapped


Array = numpy.Array  # `jaxtyping.Array` resolves to `jax.Array`
