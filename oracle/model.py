"""Oracle restatement of train/model.py:14-136 (CPU torch, TEST INFRASTRUCTURE ONLY)."""
import torch
from torch import nn

from . import nn as onn
from .layers import FactoredAttention, GumbelSigmoidSTE, PatchEmbedding, PatchUnEmbedding
from .unet import UNet


class Encoder(nn.Module):
    """train/model.py:14-60."""

    def __init__(self, height, width, channels, patch_size, depth, mlp_dim, num_heads, qkv_features,
                 max_temporal_len, spatial_compression_rate, rngs, dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        max_spatial_len = height // patch_size * width // patch_size
        self.last_dim = channels * patch_size * patch_size
        self.patch_embedding = PatchEmbedding(height, width, channels, patch_size, rngs, dtype, param_dtype)
        lat = self.last_dim // spatial_compression_rate
        self.spatial_compression = onn.Linear(self.last_dim, lat, rngs, dtype, param_dtype)
        self.variance_estimator = onn.Linear(self.last_dim, lat, rngs, dtype, param_dtype)
        self.selection_layer1 = onn.Linear(lat, 1, rngs, dtype, param_dtype)
        self.selection_layer2 = onn.Linear(max_spatial_len, 1, rngs, dtype, param_dtype)
        self.gumbel_sigmoid = GumbelSigmoidSTE(temperature=1.0)
        self.layers = nn.ModuleList(
            FactoredAttention(mlp_dim, self.last_dim, num_heads, qkv_features, max_temporal_len, max_spatial_len,
                              rngs, dtype, param_dtype) for _ in range(depth))

    def forward(self, x, mask, rngs, train=True, gumbel_u=None):
        x = self.patch_embedding(x)
        for layer in self.layers:
            x = layer(x, mask)
        mean = self.spatial_compression(x)
        variance = onn.softplus(self.variance_estimator(x))
        log_variance = torch.log(variance)
        sel = self.selection_layer1(mean).squeeze(-1)                      # b t hw
        selection = self.gumbel_sigmoid(self.selection_layer2(sel) + 1, rngs, train=train, u=gumbel_u)
        return mean, log_variance, selection.unsqueeze(-1)                # selection: b t 1 1


class Decoder(nn.Module):
    """train/model.py:62-97."""

    def __init__(self, height, width, channels, patch_size, depth, mlp_dim, num_heads, qkv_features,
                 max_temporal_len, spatial_compression_rate, unembedding_upsample_rate, rngs,
                 dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.last_dim = channels * patch_size * patch_size
        self.patch_unembedding = PatchUnEmbedding(height, width, channels, patch_size, unembedding_upsample_rate,
                                                  rngs, dtype, param_dtype)
        self.spatial_decompression = onn.Linear(self.last_dim // spatial_compression_rate, self.last_dim,
                                                rngs, dtype, param_dtype)
        max_spatial_len = height // patch_size * width // patch_size
        self.layers = nn.ModuleList(
            FactoredAttention(mlp_dim, self.last_dim, num_heads, qkv_features, max_temporal_len, max_spatial_len,
                              rngs, dtype, param_dtype) for _ in range(depth))
        self.unet = UNet(channels=channels * unembedding_upsample_rate, base_features=16, num_levels=3,
                         out_features=channels, rngs=rngs, dtype=dtype, param_dtype=param_dtype)

    def forward(self, x, mask, rngs, train=True):
        x = self.spatial_decompression(x)
        for layer in self.layers:
            x = layer(x, mask)
        feats, x = self.patch_unembedding(x)
        return x + self.unet(feats)


class VideoVAE(nn.Module):
    """train/model.py:101-136.  ``noise`` / ``gumbel_u`` optionally inject the random draws."""

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

    def forward(self, x, mask, rngs, train=True, noise=None, gumbel_u=None):
        mean, log_variance, selection = self.encoder(x, mask, rngs, train=train, gumbel_u=gumbel_u)
        if train:
            if noise is None:
                noise = torch.randn(log_variance.shape, generator=rngs.sampling(), dtype=torch.float32)
            std = torch.exp(log_variance / 2)
            sampled_latent = mean + noise.to(torch.promote_types(mean.dtype, torch.float32)) * std
        else:
            sampled_latent = mean
        compressed = self.fill_token * (1 - selection) + sampled_latent * selection
        reconstruction = self.decoder(compressed, mask, rngs, train=train)
        return reconstruction, compressed, selection, log_variance, mean
