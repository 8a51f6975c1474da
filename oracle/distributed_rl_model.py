"""Oracle restatement of claude_distributed/rl_model.py (CPU torch, TEST INFRASTRUCTURE ONLY).

The data-parallel trainer's copy of the RL model differs from train/rl_model.py in one thing: the encoder returns the
VARIANCE ``softplus(variance_estimator(x))`` (claude_distributed/rl_model.py:55-60), the latent is sampled as
``mean + noise * sqrt(variance)`` (:125-128) and the 5th element of the 6-tuple is that variance (:147), not its log.
"""
import torch
from torch import nn

from . import nn as onn
from .model import Decoder, Encoder as _Encoder


class Encoder(_Encoder):
    """claude_distributed/rl_model.py:14-60."""

    def forward(self, x, mask, rngs, train=True):
        x = self.patch_embedding(x)
        for layer in self.layers:
            x = layer(x, mask)
        mean = self.spatial_compression(x)
        variance = onn.softplus(self.variance_estimator(x))
        sel = self.selection_layer1(mean).squeeze(-1)                      # b t hw
        selection = torch.sigmoid(self.selection_layer2(sel) + 1)          # b t 1
        return mean, variance, selection


class VideoVAE(nn.Module):
    """claude_distributed/rl_model.py:103-147."""

    def __init__(self, height, width, channels, patch_size, encoder_depth, decoder_depth, mlp_dim, num_heads,
                 qkv_features, max_temporal_len, spatial_compression_rate, unembedding_upsample_rate, rngs,
                 dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        key = rngs.sampling()
        self.encoder = Encoder(height, width, channels, patch_size, encoder_depth, mlp_dim, num_heads, qkv_features,
                               max_temporal_len, spatial_compression_rate, rngs, dtype, param_dtype)
        self.decoder = Decoder(height, width, channels, patch_size, decoder_depth, mlp_dim, num_heads, qkv_features,
                               max_temporal_len, spatial_compression_rate, unembedding_upsample_rate, rngs,
                               dtype, param_dtype)
        lat = channels * patch_size * patch_size // spatial_compression_rate
        self.fill_token = nn.Parameter(torch.randn(1, 1, 1, lat, generator=key, dtype=param_dtype) * 0.02)

    def forward(self, x, mask, rngs, train=True, noise=None, bernoulli_u=None):
        mean, variance, selection = self.encoder(x, mask, rngs, train=train)
        if train:
            if noise is None:
                noise = torch.randn(variance.shape, generator=rngs.sampling(), dtype=torch.float32)
            sampled_latent = mean + noise.to(torch.promote_types(mean.dtype, torch.float32)) * torch.sqrt(variance)
        else:
            sampled_latent = mean
        rep = lambda a: a.repeat_interleave(2, dim=0)                      # noqa: E731  'b ... -> (b 2) ...'
        selection = rep(selection)[..., None]                              # (b 2) t 1 1
        sampled_latent, mean, variance, mask = rep(sampled_latent), rep(mean), rep(variance), rep(mask)
        if bernoulli_u is None:
            bernoulli_u = torch.rand(selection.shape, generator=rngs.sampling(), dtype=torch.float32)
        selection_mask = (bernoulli_u.reshape(selection.shape) < selection).to(sampled_latent.dtype)   # no gradient
        compressed = self.fill_token * (1 - selection_mask) + sampled_latent * selection_mask
        reconstruction = self.decoder(compressed, mask, rngs, train=train)
        return reconstruction, compressed, selection, selection_mask, variance, mean
