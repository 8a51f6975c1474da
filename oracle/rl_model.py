"""Oracle restatement of train/rl_model.py (CPU torch, TEST INFRASTRUCTURE ONLY): the RL variant of the VAE.

Differences from train/model.py: the encoder's frame gate is a plain probability ``sigmoid(selection_layer2(.) + 1)``
of shape (b, t, 1) (rl_model.py:56-60), and ``VideoVAE.__call__`` (rl_model.py:119-147) duplicates every sample
(``repeat 'b ... -> (b 2) ...'``), draws a Bernoulli keep-mask per frame from that probability
(``jax.random.bernoulli`` = ``uniform < p``) and returns the 6-tuple
(reconstruction, compressed_representation, selection, selection_mask, log_variance, mean).
``noise`` / ``bernoulli_u`` inject the random draws for parity tests.
"""
import torch
from torch import nn

from . import nn as onn
from .model import Decoder, Encoder as _Encoder


class Encoder(_Encoder):
    """train/rl_model.py:14-60."""

    def forward(self, x, mask, rngs, train=True):
        x = self.patch_embedding(x)
        for layer in self.layers:
            x = layer(x, mask)
        mean = self.spatial_compression(x)
        variance = onn.softplus(self.variance_estimator(x))
        log_variance = torch.log(variance)
        sel = self.selection_layer1(mean).squeeze(-1)                      # b t hw
        selection = torch.sigmoid(self.selection_layer2(sel) + 1)          # b t 1   (rl_model.py:59)
        return mean, log_variance, selection


class VideoVAE(nn.Module):
    """train/rl_model.py:100-147."""

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
        mean, log_variance, selection = self.encoder(x, mask, rngs, train=train)
        if train:
            if noise is None:
                noise = torch.randn(log_variance.shape, generator=rngs.sampling(), dtype=torch.float32)
            sampled_latent = mean + noise.to(torch.promote_types(mean.dtype, torch.float32)) * torch.exp(log_variance / 2)
        else:
            sampled_latent = mean
        rep = lambda a: a.repeat_interleave(2, dim=0)                      # noqa: E731  'b ... -> (b 2) ...'
        selection = rep(selection)[..., None]                              # (b 2) t 1 1
        sampled_latent, mean, log_variance, mask = rep(sampled_latent), rep(mean), rep(log_variance), rep(mask)
        if bernoulli_u is None:
            bernoulli_u = torch.rand(selection.shape, generator=rngs.sampling(), dtype=torch.float32)
        selection_mask = (bernoulli_u.reshape(selection.shape) < selection).to(sampled_latent.dtype)   # no gradient
        compressed = self.fill_token * (1 - selection_mask) + sampled_latent * selection_mask
        reconstruction = self.decoder(compressed, mask, rngs, train=train)
        return reconstruction, compressed, selection, selection_mask, log_variance, mean
