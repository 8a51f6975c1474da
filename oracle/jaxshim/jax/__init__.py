"""`jax` stand-in backed by CPU torch -- see oracle/jaxshim/README.md.  TEST INFRASTRUCTURE ONLY."""
import torch as _torch

from . import lax, nn, random, sharding, tree_util  # noqa: F401
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


class _Device:
    platform, id, process_index = "cpu", 0, 0

    def __repr__(self):
        return "ShimCpuDevice(id=0)"


def devices(kind=None):
    return [_Device()]


def device_count():
    return 1


local_device_count = device_count


def process_index():
    return 0


def process_count():
    return 1


def device_put(x, device=None):
    """One host device: every placement (replicated or batch-sharded over a one-device mesh) is the value itself."""
    return x


def block_until_ready(x):
    return x


def value_and_grad(fn, argnums=0, has_aux=False):
    """jax.value_and_grad of a function of arrays (the reference's scripts only differentiate w.r.t. argument 0)."""
    assert argnums == 0

    def wrapped(x, *args, **kwargs):
        xv = _torch.as_tensor(x).detach().clone().requires_grad_(True).as_subclass(Array)
        out = fn(xv, *args, **kwargs)
        val, aux = out if has_aux else (out, None)
        (g,) = _torch.autograd.grad(val, [xv], allow_unused=True)
        g = (g if g is not None else _torch.zeros_like(xv)).detach().as_subclass(Array)
        return ((val.detach(), aux), g) if has_aux else (val.detach(), g)
    return wrapped


def grad(fn, argnums=0, has_aux=False):
    vg = value_and_grad(fn, argnums, has_aux)

    def wrapped(*a, **k):
        out, g = vg(*a, **k)
        return (g, out[1]) if has_aux else g
    return wrapped


class _Debug:
    @staticmethod
    def print(fmt, *a, **k):
        print(fmt.format(*[float(v) for v in a], **{n: float(v) for n, v in k.items()}))


debug = _Debug()
