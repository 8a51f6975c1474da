"""jax.random subset.  Draws are NOT bit-compatible with threefry (they need not be: the fixture generator records every
draw the model makes and the consumers replay them); shapes, dtypes and ranges follow jax ([0,1) uniform, N(0,1))."""
import hashlib

import torch

from ._core import Array, to_dtype


class Key:
    def __init__(self, words):
        self.words = tuple(int(w) for w in words)

    def _generator(self):
        h = hashlib.sha256(repr(self.words).encode()).digest()
        return torch.Generator().manual_seed(int.from_bytes(h[:8], "little") >> 1)


def key(seed):
    return Key((int(seed),))


PRNGKey = key


def split(k, num=2):
    return [Key(k.words + (i,)) for i in range(num)]


def fold_in(k, data):
    return Key(k.words + (-1, int(data)))


def uniform(key, shape=(), dtype=torch.float32, minval=0.0, maxval=1.0):
    dt = to_dtype(dtype)
    u = torch.rand(tuple(shape), generator=key._generator(), dtype=torch.float32)
    return (u * (maxval - minval) + minval).to(dt).as_subclass(Array)


def normal(key, shape=(), dtype=torch.float32):
    return torch.randn(tuple(shape), generator=key._generator(), dtype=torch.float32).to(to_dtype(dtype)).as_subclass(Array)


def truncated_normal(key, lower, upper, shape=(), dtype=torch.float32):
    t = torch.empty(tuple(shape), dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, 0.0, 1.0, lower, upper, generator=key._generator())
    return t.to(to_dtype(dtype)).as_subclass(Array)


def bernoulli(key, p=0.5, shape=None):
    """jax.random.bernoulli: `uniform(key, shape) < p` (jax/_src/random.py::_bernoulli).  Looks `uniform` up in the module
    at call time so that a recorder patched over jax.random.uniform sees the draw."""
    import sys
    p = p if isinstance(p, torch.Tensor) else torch.as_tensor(p, dtype=torch.float32)
    shape = tuple(p.shape) if shape is None else tuple(shape)
    u = sys.modules[__name__].uniform(key, shape, p.dtype if p.is_floating_point() else torch.float32)
    return (u < p).as_subclass(Array)
