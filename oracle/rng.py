"""Minimal stand-in for flax.nnx.Rngs (test infrastructure, CPU only).

The reference threads an ``nnx.Rngs`` through constructors and calls and draws
keys with ``rngs.sampling()`` (train/layers.py:244, train/model.py:106,125) or
implicitly for parameter init.  JAX threefry streams cannot be reproduced
without JAX, so every draw here is a fresh ``torch.Generator`` seeded with
(seed, stream-counter); parity tests inject the noise tensors explicitly.
"""
import torch


class Rngs:
    def __init__(self, seed: int = 0):
        self.seed = int(seed)
        self._count = 0

    def _next(self) -> torch.Generator:
        g = torch.Generator(device="cpu")
        g.manual_seed((self.seed * 1000003 + self._count * 7919 + 12345) & 0x7FFFFFFFFFFFFFFF)
        self._count += 1
        return g

    def sampling(self) -> torch.Generator:
        return self._next()

    def params(self) -> torch.Generator:
        return self._next()

    __call__ = sampling
