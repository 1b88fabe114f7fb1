"""numpy stand-in for the `equinox` names the reference's layers use (see ../README.md).  Build container only.

`Module`: Equinox modules are dataclasses whose fields are the class annotations, `field(init=False)` ones being filled in by
`__post_init__` (which also receives the `InitVar` arguments).  Restated with the standard library's dataclass machinery; the
shim's modules are not frozen, so the generator may swap a sub-module for a recorder."""
import dataclasses as _dc

from . import nn  # noqa: F401


def field(*, init=True, default=_dc.MISSING, default_factory=_dc.MISSING, static=False, **kw):
    if default is not _dc.MISSING:
        return _dc.field(init=init, default=default)
    if default_factory is not _dc.MISSING:
        return _dc.field(init=init, default_factory=default_factory)
    return _dc.field(init=init)


class Module:
    def __init_subclass__(cls, strict=False, **kw):
        super().__init_subclass__(**kw)
        _dc.dataclass(cls, eq=False, repr=False)


def tree_inference(m, value=True):
    return m
