"""flax.nnx subset on CPU torch, restated from flax 0.12's published layer algorithms (flax/nnx/nn/linear.py,
normalization.py, flax/linen/pooling.py) independently of oracle/nn.py.  TEST INFRASTRUCTURE ONLY.

  Linear          promote_dtype(inputs, kernel, bias -> dtype); y = x . kernel + bias; kernel (in, out), lecun_normal
  LayerNorm       eps 1e-6, statistics in promote(fp32, x.dtype), use_fast_variance: var = max(0, E[x^2] - E[x]^2),
                  y = (x - mean) * (rsqrt(var + eps) * scale) + bias, cast to dtype
  GroupNorm       x -> [..., G, C/G], reduce over every non-batch axis and C/G (same statistics rule), per-channel affine
  Conv            promote_dtype, lax.conv_general_dilated cross-correlation, channels last, kernel (*k, in, out), 'SAME'
  ConvTranspose   promote_dtype, lax.conv_transpose(transpose_kernel=False), padding 'SAME'
  max_pool        reduce_window(max, -inf), 'VALID'
"""
import math

import torch

import jax
import jax.numpy as jnp
from jax import lax
from jax._core import Array, asarray, to_dtype


# ------------------------------------------------------------------------------------------------ variables / modules
class Variable:
    def __init__(self, value, **metadata):
        self.value = asarray(value.value if isinstance(value, Variable) else value)
        self.metadata = metadata

    # flax Variables proxy arithmetic to their value (the reference writes `self.fill_token * (1 - selection)`)
    def __jax_array__(self):
        return self.value

    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return self.value.dtype

    def __getitem__(self, idx):
        return self.value[idx]

    def __mul__(self, o):
        return self.value * _raw(o)

    def __rmul__(self, o):
        return _raw(o) * self.value

    def __add__(self, o):
        return self.value + _raw(o)

    def __radd__(self, o):
        return _raw(o) + self.value

    def __sub__(self, o):
        return self.value - _raw(o)

    def __rsub__(self, o):
        return _raw(o) - self.value

    def __neg__(self):
        return -self.value

    def __matmul__(self, o):
        return self.value @ _raw(o)


def _raw(v):
    return v.value if isinstance(v, Variable) else v


class Param(Variable):
    pass


class BatchStat(Variable):
    pass


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__()

    def __init__(self, *a, **k):
        pass


def _walk(node, path, out):
    """Depth-first walk of the attribute graph: Modules, lists / tuples / dicts of them, Variables."""
    if isinstance(node, Variable):
        out.append((path, node))
    elif isinstance(node, Module):
        for name in sorted(vars(node)):
            _walk(vars(node)[name], path + (name,), out)
    elif isinstance(node, (list, tuple)):
        for i, v in enumerate(node):
            _walk(v, path + (i,), out)
    elif isinstance(node, dict):
        for k in node:
            _walk(node[k], path + (k,), out)


def iter_variables(module, *filters):
    out = []
    _walk(module, (), out)
    if filters:
        out = [(p, v) for p, v in out if any(isinstance(v, f) for f in filters if isinstance(f, type))]
    return out


class State:
    """Nested mapping path -> Variable value (flax State.to_pure_dict(): list indices become integer keys)."""

    def __init__(self, flat):
        self.flat = dict(flat)

    def to_pure_dict(self):
        root = {}
        for path, v in self.flat.items():
            d = root
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = v
        return root

    def flat_state(self):
        return dict(self.flat)


def state(module, *filters):
    if isinstance(module, State):           # nnx.state(grads): a State is its own state
        return module
    if isinstance(module, Optimizer):
        flat = {("step",): torch.tensor(module.step, dtype=torch.int32)}
        _opt_state_flat(module.opt_state, ("opt_state",), flat)
        for p, v in iter_variables(module.model, *filters):
            flat[("model",) + p] = v.value.detach()
        return State(flat)
    return State({p: v.value.detach() for p, v in iter_variables(module, *filters)})


def split(module, *filters):
    """nnx.split -> (graphdef, state).  The shim's graphdef is the module object itself: merge() writes the state back
    into it, which is all the reference's scripts do with the pair (device_put the state, merge, call)."""
    return module, state(module, *filters)


def merge(graphdef, st, *more):
    update(graphdef, st)
    for m in more:
        update(graphdef, m)
    return graphdef


def _flatten_pure(d, path, out):
    for k, v in d.items():
        if isinstance(v, dict):
            _flatten_pure(v, path + (k,), out)
        else:
            out[path + (k,)] = v
    return out


def update(module, st):
    """nnx.update(module, state): accepts a State or the nested dict of State.to_pure_dict()."""
    flat = st.flat if isinstance(st, State) else _flatten_pure(st, (), {})
    for p, v in iter_variables(module):
        if p in flat:
            v.value = asarray(flat[p])


def value_and_grad(fn, argnums=0, has_aux=False, wrt=Param):
    """nnx.value_and_grad: differentiates with respect to the Param state of the module passed as argument `argnums`."""
    assert argnums == 0

    def wrapped(module, *args, **kwargs):
        variables = iter_variables(module, wrt)
        for _, v in variables:
            v.value = v.value.detach().clone().requires_grad_(True)
        try:
            out = fn(module, *args, **kwargs)
            loss, aux = out if has_aux else (out, None)
            grads = torch.autograd.grad(loss, [v.value for _, v in variables], allow_unused=True)
        finally:
            for _, v in variables:
                v.value = v.value.detach()
        gstate = State({p: (g.detach() if g is not None else torch.zeros_like(v.value)).as_subclass(Array)
                        for (p, v), g in zip(variables, grads)})
        det = lambda t: t.detach() if isinstance(t, torch.Tensor) else t                       # noqa: E731
        loss = loss.detach()
        if has_aux:
            return (loss, jax.tree_util.tree_map(det, aux)), gstate
        return loss, gstate
    return wrapped


def grad(fn, argnums=0, has_aux=False, wrt=Param):
    vg = value_and_grad(fn, argnums, has_aux, wrt)

    def wrapped(*a, **k):
        out, g = vg(*a, **k)
        return (g, out[1]) if has_aux else g
    return wrapped


def remat(fn=None, **kw):                 # rematerialisation changes memory, not values
    if fn is None:
        return lambda f: f
    return fn


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


class Optimizer(Module):
    """nnx.Optimizer(model, tx): ``update(grads)`` (flax 0.10, what the reference calls: train/rl_nonadversarial.py:196) or
    ``update(model, grads)`` (flax 0.11+).  ``nnx.state(optimizer)`` = {"step", "opt_state": nested tuples by index} next to
    the model's own variables under "model", the nesting a checkpoint of the reference has."""

    def __init__(self, model, tx, wrt=Param):
        self.model, self.tx, self.wrt = model, tx, wrt
        self.step = 0
        self.opt_state = tx.init(state(model, wrt))

    def update(self, *args, **kwargs):
        grads = args[-1] if args else kwargs["grads"]
        params = state(self.model, self.wrt)
        updates, self.opt_state = self.tx.update(grads, self.opt_state, params)
        new = {p: (params.flat[p] + u).to(params.flat[p].dtype) for p, u in updates.flat.items()}
        update(self.model, State(new))
        self.step += 1


def _opt_state_flat(node, path, out):
    if isinstance(node, State):
        for p, v in node.flat.items():
            out[path + p] = v
    elif isinstance(node, tuple) and hasattr(node, "_fields"):
        for name in node._fields:
            _opt_state_flat(getattr(node, name), path + (name,), out)
    elif isinstance(node, (tuple, list)):
        for i, v in enumerate(node):
            _opt_state_flat(v, path + (i,), out)
    elif isinstance(node, torch.Tensor):
        out[path] = node


# ------------------------------------------------------------------------------------------------ rngs
class _Stream:
    def __init__(self, seed, name):
        self.key, self.count = jax.random.fold_in(jax.random.key(seed), sum(map(ord, name))), 0

    def __call__(self):
        k = jax.random.fold_in(self.key, self.count)
        self.count += 1
        return k


class Rngs:
    """nnx.Rngs(seed): every stream name (`params`, `sampling`, ...) falls back to the default stream's seed."""

    def __init__(self, default=0, **streams):
        self._seed = default
        self._streams = {n: _Stream(s, n) for n, s in streams.items()}

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name not in self._streams:
            self._streams[name] = _Stream(self._seed, name)
        return self._streams[name]

    def __call__(self):
        return self.default()


# ------------------------------------------------------------------------------------------------ initializers
class initializers:
    @staticmethod
    def zeros(key, shape, dtype=jnp.float32):
        return jnp.zeros(shape, dtype)

    zeros_init = staticmethod(lambda: initializers.zeros)

    @staticmethod
    def ones(key, shape, dtype=jnp.float32):
        return jnp.ones(shape, dtype)

    ones_init = staticmethod(lambda: initializers.ones)

    @staticmethod
    def variance_scaling(scale, mode, distribution, in_axis=-2, out_axis=-1):
        def init(key, shape, dtype=jnp.float32):
            receptive = math.prod(shape) / shape[in_axis] / shape[out_axis]
            fan_in, fan_out = shape[in_axis] * receptive, shape[out_axis] * receptive
            denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
            variance = scale / denom
            if distribution == "truncated_normal":
                stddev = math.sqrt(variance) / 0.87962566103423978
                return (jax.random.truncated_normal(key, -2.0, 2.0, shape, jnp.float32) * stddev).astype(dtype)
            if distribution == "normal":
                return (jax.random.normal(key, shape, jnp.float32) * math.sqrt(variance)).astype(dtype)
            raise NotImplementedError(distribution)
        return init

    @staticmethod
    def lecun_normal(in_axis=-2, out_axis=-1):
        return initializers.variance_scaling(1.0, "fan_in", "truncated_normal", in_axis, out_axis)


def _promote(dtype, *arrays):
    """flax promote_dtype: cast every non-None array to `dtype` (or, with dtype None, to their common type)."""
    dtype = to_dtype(dtype)
    present = [asarray(a) for a in arrays if a is not None]
    if dtype is None:
        dtype = present[0].dtype
        for a in present[1:]:
            dtype = torch.promote_types(dtype, a.dtype)
    return [None if a is None else asarray(a).to(dtype) for a in arrays]


# ------------------------------------------------------------------------------------------------ layers
class Linear(Module):
    def __init__(self, in_features, out_features, *, use_bias=True, dtype=None, param_dtype=jnp.float32,
                 kernel_init=None, bias_init=None, rngs, precision=None):
        kernel_init = kernel_init or initializers.lecun_normal()
        bias_init = bias_init or initializers.zeros
        self.kernel = Param(kernel_init(rngs.params(), (in_features, out_features), param_dtype))
        self.bias = Param(bias_init(rngs.params(), (out_features,), param_dtype)) if use_bias else None
        self.in_features, self.out_features, self.use_bias, self.dtype = in_features, out_features, use_bias, dtype

    def __call__(self, inputs):
        x, kernel, bias = _promote(self.dtype, inputs, self.kernel.value, self.bias.value if self.bias is not None else None)
        y = torch.matmul(x, kernel)                       # dot_general over the last axis of x and the first of kernel
        if bias is not None:
            y = y + bias.reshape((1,) * (y.ndim - 1) + (-1,))
        return y


def _compute_stats(x, axes, use_fast_variance=True):
    x = x.to(torch.promote_types(torch.float32, x.dtype))
    mean = x.mean(dim=axes)
    if use_fast_variance:
        mean2 = (x * x).mean(dim=axes)
        var = torch.clamp(mean2 - mean * mean, min=0.0)
    else:
        var = ((x - mean.reshape([1 if i in axes else n for i, n in enumerate(x.shape)])) ** 2).mean(dim=axes)
    return mean, var


def _normalize(x, mean, var, reduction_axes, scale, bias, dtype, epsilon):
    """flax/nnx/nn/normalization.py::_normalize with feature_axes = (-1,)."""
    for ax in sorted(reduction_axes):
        mean, var = mean.unsqueeze(ax), var.unsqueeze(ax)
    y = x - mean
    mul = torch.rsqrt(var + epsilon)
    args = [x]
    if scale is not None:
        mul = mul * scale
        args.append(scale)
    y = y * mul
    if bias is not None:
        y = y + bias
        args.append(bias)
    out_dtype = to_dtype(dtype)
    if out_dtype is None:
        out_dtype = args[0].dtype
        for a in args[1:]:
            out_dtype = torch.promote_types(out_dtype, a.dtype)
    return y.to(out_dtype)


class LayerNorm(Module):
    def __init__(self, num_features, *, epsilon=1e-6, dtype=None, param_dtype=jnp.float32, use_bias=True, use_scale=True,
                 reduction_axes=-1, feature_axes=-1, use_fast_variance=True, rngs):
        assert reduction_axes == -1 and feature_axes == -1
        self.scale = Param(jnp.ones((num_features,), param_dtype)) if use_scale else None
        self.bias = Param(jnp.zeros((num_features,), param_dtype)) if use_bias else None
        self.epsilon, self.dtype, self.use_fast_variance = epsilon, dtype, use_fast_variance

    def __call__(self, x):
        x = asarray(x)
        axes = (x.ndim - 1,)
        mean, var = _compute_stats(x, axes, self.use_fast_variance)
        return _normalize(x, mean, var, axes, self.scale.value if self.scale is not None else None,
                          self.bias.value if self.bias is not None else None, self.dtype, self.epsilon)


class GroupNorm(Module):
    def __init__(self, num_features, num_groups=32, group_size=None, *, epsilon=1e-6, dtype=None, param_dtype=jnp.float32,
                 use_bias=True, use_scale=True, use_fast_variance=True, rngs):
        assert (num_groups is None) != (group_size is None) or group_size is None
        if num_groups is None:
            num_groups = num_features // group_size
        assert num_features % num_groups == 0
        self.num_groups, self.group_size, self.num_features = num_groups, num_features // num_groups, num_features
        self.scale = Param(jnp.ones((num_features,), param_dtype)) if use_scale else None
        self.bias = Param(jnp.zeros((num_features,), param_dtype)) if use_bias else None
        self.epsilon, self.dtype, self.use_fast_variance = epsilon, dtype, use_fast_variance

    def __call__(self, x):
        x = asarray(x)
        group_shape = tuple(x.shape[:-1]) + (self.num_groups, self.group_size)
        xg = x.reshape(group_shape)
        reduction_axes = tuple(range(1, x.ndim - 1)) + (xg.ndim - 1,)          # every non-batch axis and the in-group one
        mean, var = _compute_stats(xg, reduction_axes, self.use_fast_variance)   # [N, G]
        mean = torch.repeat_interleave(mean, self.group_size, dim=-1)            # per channel, [N, C]
        var = torch.repeat_interleave(var, self.group_size, dim=-1)
        return _normalize(x, mean, var, tuple(range(1, x.ndim - 1)), self.scale.value if self.scale is not None else None,
                          self.bias.value if self.bias is not None else None, self.dtype, self.epsilon)


def _tuple(v, n):
    return (v,) * n if isinstance(v, int) else tuple(v)


class Conv(Module):
    def __init__(self, in_features, out_features, kernel_size, strides=1, *, padding="SAME", use_bias=True, dtype=None,
                 param_dtype=jnp.float32, kernel_init=None, bias_init=None, rngs, precision=None):
        kernel_size = _tuple(kernel_size, 1)
        kernel_init = kernel_init or initializers.lecun_normal()
        bias_init = bias_init or initializers.zeros
        self.kernel = Param(kernel_init(rngs.params(), kernel_size + (in_features, out_features), param_dtype))
        self.bias = Param(bias_init(rngs.params(), (out_features,), param_dtype)) if use_bias else None
        self.kernel_size, self.strides, self.padding, self.dtype = kernel_size, strides, padding, dtype

    def __call__(self, inputs):
        x, kernel, bias = _promote(self.dtype, inputs, self.kernel.value, self.bias.value if self.bias is not None else None)
        nd = len(self.kernel_size)
        assert x.ndim == nd + 2, "batched channels-last input expected"
        y = lax.conv_general_dilated_nhwc(x, kernel, _tuple(self.strides, nd), self.padding)
        if bias is not None:
            y = y + bias.reshape((1,) * (y.ndim - 1) + (-1,))
        return y


class ConvTranspose(Module):
    def __init__(self, in_features, out_features, kernel_size, strides=None, *, padding="SAME", use_bias=True, dtype=None,
                 param_dtype=jnp.float32, kernel_init=None, bias_init=None, transpose_kernel=False, rngs, precision=None):
        kernel_size = _tuple(kernel_size, 1)
        kernel_init = kernel_init or initializers.lecun_normal()
        bias_init = bias_init or initializers.zeros
        self.kernel = Param(kernel_init(rngs.params(), kernel_size + (in_features, out_features), param_dtype))
        self.bias = Param(bias_init(rngs.params(), (out_features,), param_dtype)) if use_bias else None
        self.kernel_size, self.padding, self.dtype, self.transpose_kernel = kernel_size, padding, dtype, transpose_kernel
        self.strides = _tuple(strides if strides is not None else 1, len(kernel_size))

    def __call__(self, inputs):
        x, kernel, bias = _promote(self.dtype, inputs, self.kernel.value, self.bias.value if self.bias is not None else None)
        y = lax.conv_transpose_nhwc(x, kernel, self.strides, self.padding, transpose_kernel=self.transpose_kernel)
        if bias is not None:
            y = y + bias.reshape((1,) * (y.ndim - 1) + (-1,))
        return y


def max_pool(inputs, window_shape, strides=None, padding="VALID"):
    assert padding == "VALID"
    x = asarray(inputs)
    strides = strides or (1,) * len(window_shape)
    return lax.reduce_window_max_nhwc(x, window_shape, strides)


silu = jax.nn.silu
swish = jax.nn.silu
relu = jax.nn.relu
sigmoid = jax.nn.sigmoid
softmax = jax.nn.softmax
