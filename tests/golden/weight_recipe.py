"""Deterministic weights / clips by NAME, shared by tests/golden/make_golden_jax.py (--compact) and its consumers.

A production-depth VideoVAE has 170 M parameters; a fixture that stores them (and as many gradients) is 1.4 GB.  With
`--compact` the generator overwrites every parameter of the freshly built reference model with `param(name, shape)`
below -- a function of the Flax attribute path and the shape only, numpy's frozen legacy MT19937 stream -- so that a
consumer rebuilds the identical weights without the file carrying them, and stores of every gradient only its L2 norm,
its sum and `grad_probe` (a fixed strided sample).  Scales follow the reference's initialisers (lecun-normal kernels,
the 1e-2 variance scale of out_projection / linear2, train/layers.py:141-145,181-185) except where the reference starts
at a value that hides a code path: LayerNorm / GroupNorm scales 1 + 0.1 n, biases 0.02 n, final_conv 0.05 n (the
reference zero-initialises it, train/unet.py:144-153, which switches the U-Net's gradients off).
"""
import math
import zlib

import numpy as np

RECIPE_ID = "name-seeded-mt19937-v1"
PROBE = 96


def _rs(name, salt=0):
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (0x9E3779B1 * (salt + 1))) & 0xFFFFFFFF)


def param(name, shape):
    shape = tuple(int(s) for s in shape)
    n = _rs(name).standard_normal(shape).astype(np.float32)
    leaf = name.split(".")[-1]
    if leaf == "kernel":
        if "final_conv" in name:
            return (0.05 * n).astype(np.float32)
        fan_in = math.prod(shape[:-1])
        std = 1.0 / math.sqrt(fan_in)
        if "out_projection" in name or "linear2" in name:
            std *= 0.1
        return (std * n).astype(np.float32)
    if leaf == "scale":
        return (1.0 + 0.1 * n).astype(np.float32)
    if leaf == "bias":
        return (0.02 * n).astype(np.float32)
    if leaf == "fill_token":
        return (0.02 * n).astype(np.float32)
    raise KeyError(f"weight_recipe: no rule for parameter {name!r}")


def clip(shape, seed=11):
    """Synthetic clip in [0, 1) (the dataloader's range after /255), [b, t, h, w, c]."""
    return np.random.RandomState(seed).random_sample(tuple(shape)).astype(np.float32)


def normal(shape, seed=13):
    """The reparameterisation noise of a --compact fixture: N(0, 1) by shape (replaces the model's jax.random.normal draw,
    so that the file need not carry b * t * hw * latent floats)."""
    return np.random.RandomState(seed).standard_normal(tuple(int(v) for v in shape)).astype(np.float32)


def grad_probe(g):
    """Fixed sample of a gradient tensor: PROBE elements at an even stride over the flattened tensor."""
    flat = np.asarray(g, np.float32).reshape(-1)
    idx = np.linspace(0, flat.size - 1, num=min(PROBE, flat.size)).astype(np.int64)
    return flat[idx]
