"""`jax` stand-in backed by CPU torch -- see oracle/jaxshim/README.md.  TEST INFRASTRUCTURE ONLY."""
import torch as _torch

from . import lax, nn, random, tree_util  # noqa: F401
from . import numpy  # noqa: F401
from ._core import Array  # noqa: F401

__version__ = "0+torchshim"
IS_SHIM = True


class custom_vjp:
    """jax.custom_vjp: f.defvjp(fwd, bwd) with fwd(*args) -> (out, residuals) and bwd(residuals, g) -> tuple of cotangents."""

    def __init__(self, fun):
        self.fun, self.fwd, self.bwd = fun, None, None

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *args):
        fwd, bwd = self.fwd, self.bwd

        class _Fn(_torch.autograd.Function):
            @staticmethod
            def forward(ctx, *a):
                out, res = fwd(*a)
                ctx.res = res
                return out

            @staticmethod
            def backward(ctx, g):
                return tuple(bwd(ctx.res, g))
        return _Fn.apply(*args)


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def devices(kind=None):
    raise RuntimeError("jaxshim: CPU torch only")


class _Debug:
    @staticmethod
    def print(fmt, *a, **k):
        print(fmt.format(*[float(v) for v in a], **{n: float(v) for n, v in k.items()}))


debug = _Debug()
