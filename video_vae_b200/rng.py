"""nnx.Rngs-like key source for the B200 path.

The reference threads ``nnx.Rngs`` through constructors and calls (train/model.py:104,119; train/layers.py:242).  Here a
key is a (seed, offset) pair for the in-kernel Philox4x32-10 generator (csrc/elementwise.cu): ``sampling()`` hands out
disjoint 2^40-wide counter ranges, so every draw in a run is independent and reproducible from the seed.
"""
import torch


class Rngs:
    def __init__(self, seed: int = 0, **_streams):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._draws = 0

    def sampling(self):
        """(seed, offset) for one Philox-driven draw."""
        self._draws += 1
        return self.seed, (self._draws << 40)

    def params(self) -> torch.Generator:
        """CPU generator for parameter initialisation (not on the hot path)."""
        self._draws += 1
        g = torch.Generator(device="cpu")
        g.manual_seed((self.seed * 1000003 + self._draws * 7919 + 12345) & 0x7FFFFFFFFFFFFFFF)
        return g

    __call__ = sampling
