"""jax.numpy subset (what train/layers.py, model.py, unet.py, rl_model.py, loss_fn and the fixture generator call)."""
import builtins
import math as _math

import torch

from ._core import Array, asarray, to_dtype

float32 = torch.float32
bfloat16 = torch.bfloat16
float16 = torch.float16
int32 = torch.int32
bool_ = torch.bool
newaxis = None
pi = _math.pi
inf = float("inf")
dtype = to_dtype             # used by the reference as an annotation (`dtype: jnp.dtype = jnp.bfloat16`)


class ndarray:               # einops' JaxBackend probes isinstance(x, jnp.ndarray); nothing is: einops uses its torch backend
    pass


array = asarray


def _a(x):
    return asarray(x)


def arange(*args, dtype=None):
    floaty = builtins.any(isinstance(v, float) for v in args)
    dt = to_dtype(dtype) or (torch.float32 if floaty else torch.int32)
    return torch.arange(*args, dtype=dt).as_subclass(Array)


def zeros(shape, dtype=None):
    return torch.zeros(shape, dtype=to_dtype(dtype) or torch.float32).as_subclass(Array)


def ones(shape, dtype=None):
    return torch.ones(shape, dtype=to_dtype(dtype) or torch.float32).as_subclass(Array)


def zeros_like(x, dtype=None):
    return torch.zeros_like(_a(x), dtype=to_dtype(dtype))


def ones_like(x, dtype=None):
    return torch.ones_like(_a(x), dtype=to_dtype(dtype))


def full(shape, fill_value, dtype=None):
    return torch.full(shape, fill_value, dtype=to_dtype(dtype) or torch.float32).as_subclass(Array)


def _unary(fn):
    return lambda x: fn(_a(x))


exp = _unary(torch.exp)
log = _unary(torch.log)
sin = _unary(torch.sin)
cos = _unary(torch.cos)
sqrt = _unary(torch.sqrt)
abs = _unary(torch.abs)
square = _unary(torch.square)
tanh = _unary(torch.tanh)
isfinite = _unary(torch.isfinite)
isnan = _unary(torch.isnan)
round = _unary(torch.round)                  # half to even, as jnp.round


def logaddexp(a, b):
    a = _a(a)
    return torch.logaddexp(a, _a(b).to(a.dtype).expand_as(a))


def maximum(a, b):
    return torch.maximum(_a(a), _a(b))


def minimum(a, b):
    return torch.minimum(_a(a), _a(b))


def where(cond, a, b):
    return torch.where(_a(cond), a, b)


def clip(x, min=None, max=None, a_min=None, a_max=None):
    lo = min if min is not None else a_min
    hi = max if max is not None else a_max
    x = _a(x)
    if not x.is_floating_point() and (isinstance(lo, float) or isinstance(hi, float)):
        x = x.to(torch.float32)              # weak-typed python float promotes an integer array to float32
    return torch.clamp(x, min=lo, max=hi)


def _axis_kw(axis, keepdims):
    kw = {}
    if axis is not None:
        kw["dim"] = axis
    if keepdims:
        kw["keepdim"] = True
    return kw


def mean(x, axis=None, keepdims=False, dtype=None):
    x = _a(x)
    if not x.is_floating_point():
        x = x.to(torch.float32)
    return x.mean(**_axis_kw(axis, keepdims))


def sum(x, axis=None, keepdims=False, dtype=None):
    return _a(x).sum(**_axis_kw(axis, keepdims))


def max(x, axis=None, keepdims=False):
    x = _a(x)
    return x.amax(**_axis_kw(axis, keepdims)) if axis is not None else x.max()


def min(x, axis=None, keepdims=False):
    x = _a(x)
    return x.amin(**_axis_kw(axis, keepdims)) if axis is not None else x.min()


def any(x, axis=None, keepdims=False):        # noqa: A001  (jnp.any; this module reaches the builtin as builtins.any)
    x = _a(x).bool()
    return x.any(**_axis_kw(axis, keepdims)) if axis is not None else x.any()


def all(x, axis=None, keepdims=False):        # noqa: A001
    x = _a(x).bool()
    return x.all(**_axis_kw(axis, keepdims)) if axis is not None else x.all()


def var(x, axis=None, keepdims=False):
    return _a(x).var(unbiased=False, **_axis_kw(axis, keepdims))


def std(x, axis=None, keepdims=False):
    """jnp.std: sqrt of the population variance (ddof = 0)."""
    return torch.sqrt(var(x, axis, keepdims))


def prod(x, axis=None, keepdims=False):
    x = _a(x)
    return x.prod(**_axis_kw(axis, keepdims)) if axis is not None else x.prod()


def concatenate(arrays, axis=0):
    return torch.cat([_a(v) for v in arrays], dim=axis)


def stack(arrays, axis=0):
    return torch.stack([_a(v) for v in arrays], dim=axis)


def split(x, indices_or_sections, axis=0):
    x = _a(x)
    if isinstance(indices_or_sections, int):
        assert x.shape[axis] % indices_or_sections == 0, "jnp.split: sections must divide the axis"
    return list(torch.tensor_split(x, indices_or_sections, dim=axis))


def einsum(spec, *operands, preferred_element_type=None, precision=None):
    ops = [_a(o) for o in operands]
    if preferred_element_type is not None:
        ops = [o.to(to_dtype(preferred_element_type)) for o in ops]
    return torch.einsum(spec, *ops)


def matmul(a, b):
    return torch.matmul(_a(a), _a(b))


def reshape(x, shape):
    return _a(x).reshape(shape)


def transpose(x, axes=None):
    x = _a(x)
    return x.permute(*axes) if axes is not None else x.permute(*reversed(range(x.ndim)))


def expand_dims(x, axis):
    x = _a(x)
    for ax in sorted((axis,) if isinstance(axis, int) else axis):
        x = x.unsqueeze(ax)
    return x


def squeeze(x, axis=None):
    return _a(x).squeeze() if axis is None else _a(x).squeeze(axis)


def repeat(x, repeats, axis=None):
    return torch.repeat_interleave(_a(x), repeats, dim=axis)


def broadcast_to(x, shape):
    return _a(x).expand(shape)


def allclose(a, b, rtol=1e-5, atol=1e-8):
    return bool(torch.allclose(_a(a), _a(b), rtol=rtol, atol=atol))


def array_equal(a, b):
    return bool(torch.equal(_a(a), _a(b)))


class finfo:
    def __init__(self, dt):
        f = torch.finfo(to_dtype(dt))
        self.max, self.min, self.eps, self.tiny = f.max, f.min, f.eps, f.tiny


def promote_types(a, b):
    return torch.promote_types(to_dtype(a), to_dtype(b))
