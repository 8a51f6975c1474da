"""Array type of the shim: a torch.Tensor subclass with the handful of jax.Array methods the reference uses."""
import numpy as np
import torch


class _At:
    def __init__(self, arr):
        self.arr, self.idx = arr, None

    def __getitem__(self, idx):
        self.idx = idx
        return self

    def set(self, value):
        out = self.arr.clone()
        out[self.idx] = value
        return out

    def add(self, value):
        out = self.arr.clone()
        out[self.idx] += value
        return out


class _Size(int):
    """jax's ``x.size`` is the element count (an int); torch's is a method.  This is both: an int that can be called."""

    def __new__(cls, t):
        obj = int.__new__(cls, t.numel())
        obj._t = t
        return obj

    def __call__(self, *a, **k):
        return torch.Tensor.size(self._t, *a, **k)


class _Shard:
    def __init__(self, device, data):
        self.device, self.data, self.index = device, data, (slice(None),) * data.dim()


class Array(torch.Tensor):
    """jax.Array look-alike.  torch's default __torch_function__ keeps the subclass through every op."""

    @property
    def size(self):
        return _Size(self)

    def astype(self, dtype):
        return self.to(to_dtype(dtype))

    @property
    def at(self):
        return _At(self)

    def copy(self):
        return self.clone()

    def __array__(self, dtype=None, copy=None):
        t = self.detach().as_subclass(torch.Tensor)
        if t.dtype == torch.bfloat16:
            t = t.float()
        a = t.numpy()
        return a.astype(dtype) if dtype is not None else a

    def block_until_ready(self):
        return self

    @property
    def addressable_shards(self):
        """One host device: the array is its own single shard."""
        import jax
        return [_Shard(jax.devices()[0], self)]


_DTYPES = {bool: torch.bool, int: torch.int32, float: torch.float32, "float32": torch.float32,
           "bfloat16": torch.bfloat16, "bool": torch.bool, "int32": torch.int32}


def to_dtype(dtype):
    if dtype is None or isinstance(dtype, torch.dtype):
        return dtype
    if dtype in _DTYPES:
        return _DTYPES[dtype]
    if isinstance(dtype, np.dtype) or (isinstance(dtype, type) and issubclass(dtype, np.generic)):
        return torch.from_numpy(np.zeros((), dtype)).dtype
    raise TypeError(f"jaxshim: unsupported dtype {dtype!r}")


def asarray(x, dtype=None):
    dtype = to_dtype(dtype)
    if hasattr(x, "__jax_array__"):
        x = x.__jax_array__()
    if isinstance(x, torch.Tensor):
        t = x if dtype is None else x.to(dtype)
    else:
        if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in x):
            t = torch.stack([torch.as_tensor(v) for v in x])
        else:
            a = np.asarray(x)
            if a.dtype == np.float64:
                a = a.astype(np.float32)          # jax default: x64 disabled
            elif a.dtype == np.int64:
                a = a.astype(np.int32)
            t = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None:
            t = t.to(dtype)
    return t if isinstance(t, Array) else t.as_subclass(Array)
