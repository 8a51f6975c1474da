"""B200-native drop-in for claude_distributed/rl_model.py: the data-parallel trainer's copy of the RL model.

Same kernels and weights as ``rl_model.py``; the one difference is the interface (claude_distributed/rl_model.py:55-60,
125-128, 147): the encoder hands back the VARIANCE ``softplus(variance_estimator(x))`` and the 5th element of the call's
6-tuple is that variance instead of its logarithm.  The CUDA head kernel produces ``log(softplus(.))`` (what the
reparameterisation and the KL term consume); the variance is its exponential, taken once on the doubled batch, and the
gradient a caller puts on the variance flows back through that exponential into the head's backward kernel.
"""
import torch

from . import rl_model as _rl


class Encoder(_rl.Encoder):
    """claude_distributed/rl_model.py:14-60.  Returns (mean, variance, selection[b,t,1])."""

    def forward(self, x, mask, rngs, train=True):
        mean, log_variance, selection = super().forward(x, mask, rngs, train=train)
        return mean, torch.exp(log_variance), selection


class VideoVAE(_rl.VideoVAE):
    """claude_distributed/rl_model.py:103-147.  ``self.encoder`` is the log-variance encoder of ``rl_model`` (the
    kernels want the logarithm); ``encode`` gives the reference encoder's (mean, variance, selection)."""

    def encode(self, x, mask, rngs, train=True):
        mean, log_variance, selection = self.encoder(x, mask, rngs, train=train)
        return mean, torch.exp(log_variance), selection

    def forward(self, x, mask, rngs, train=True, noise=None, bernoulli_u=None):
        reconstruction, compressed, selection, selection_mask, log_variance, mean = super().forward(
            x, mask, rngs, train=train, noise=noise, bernoulli_u=bernoulli_u)
        return reconstruction, compressed, selection, selection_mask, torch.exp(log_variance), mean
