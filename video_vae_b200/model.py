"""B200-native drop-in for the reference's train/model.py: Encoder, Decoder, VideoVAE.

Constructor arguments, call signatures, return tuples and parameter names follow the reference
(train/model.py:14-136).  ``rngs`` is a ``video_vae_b200.Rngs`` (seed + Philox offsets).  Extra keyword-only
arguments (``noise``, ``gumbel_u``) inject the random draws for parity tests; the defaults draw them inside the
kernels.
"""
import torch
from torch import nn

from . import functional as F_
from .layers import (FactoredAttention, GumbelSigmoidSTE, Linear, PatchEmbedding, PatchUnEmbedding, _default_device)
from .unet import UNet


class Encoder(nn.Module):
    """train/model.py:14-60.  Returns (mean, log_variance, selection[b,t,1,1])."""

    def __init__(self, height, width, channels, patch_size, depth, mlp_dim, num_heads, qkv_features, max_temporal_len,
                 spatial_compression_rate, rngs, dtype=torch.bfloat16, param_dtype=torch.float32, device=None):
        super().__init__()
        self.dtype = dtype
        max_spatial_len = height // patch_size * width // patch_size
        self.last_dim = channels * patch_size * patch_size
        self.patch_embedding = PatchEmbedding(height, width, channels, patch_size, rngs, dtype, param_dtype, device=device)
        lat = self.last_dim // spatial_compression_rate
        self.spatial_compression = Linear(self.last_dim, lat, rngs, dtype, param_dtype, device=device)
        self.variance_estimator = Linear(self.last_dim, lat, rngs, dtype, param_dtype, device=device)
        self.selection_layer1 = Linear(lat, 1, rngs, dtype, param_dtype, device=device)
        self.selection_layer2 = Linear(max_spatial_len, 1, rngs, dtype, param_dtype, device=device)
        self.gumbel_sigmoid = GumbelSigmoidSTE(temperature=1.0)
        self.layers = nn.ModuleList(
            FactoredAttention(mlp_dim, self.last_dim, num_heads, qkv_features, max_temporal_len, max_spatial_len, rngs,
                              dtype, param_dtype, device=device) for _ in range(depth))

    def forward(self, x, mask, rngs, train=True, gumbel_u=None):
        x = self.patch_embedding(x)
        b, t, hw, _ = x.shape
        tmask = FactoredAttention.temporal_mask_arg(mask, b, t, hw)      # normalise once, share across layers
        for layer in self.layers:
            x = layer(x, tmask)
        seed, offset = rngs.sampling() if (train and gumbel_u is None) else (0, 0)
        u = gumbel_u.reshape(-1).to(torch.float32).contiguous() if gumbel_u is not None else None
        return F_.EncoderHeadFn.apply(
            x, self.dtype, bool(train), float(self.gumbel_sigmoid.temperature), u, seed, offset,
            self.spatial_compression.kernel, self.spatial_compression.bias, self.variance_estimator.kernel,
            self.variance_estimator.bias, self.selection_layer1.kernel, self.selection_layer1.bias,
            self.selection_layer2.kernel, self.selection_layer2.bias)


class Decoder(nn.Module):
    """train/model.py:62-97."""

    def __init__(self, height, width, channels, patch_size, depth, mlp_dim, num_heads, qkv_features, max_temporal_len,
                 spatial_compression_rate, unembedding_upsample_rate, rngs, dtype=torch.bfloat16,
                 param_dtype=torch.float32, device=None):
        super().__init__()
        self.dtype = dtype
        self.last_dim = channels * patch_size * patch_size
        self.patch_unembedding = PatchUnEmbedding(height, width, channels, patch_size, unembedding_upsample_rate, rngs,
                                                  dtype, param_dtype, device=device)
        self.spatial_decompression = Linear(self.last_dim // spatial_compression_rate, self.last_dim, rngs, dtype,
                                            param_dtype, device=device)
        max_spatial_len = height // patch_size * width // patch_size
        self.layers = nn.ModuleList(
            FactoredAttention(mlp_dim, self.last_dim, num_heads, qkv_features, max_temporal_len, max_spatial_len, rngs,
                              dtype, param_dtype, device=device) for _ in range(depth))
        self.unet = UNet(channels=channels * unembedding_upsample_rate, base_features=16, num_levels=3,
                         out_features=channels, rngs=rngs, dtype=dtype, param_dtype=param_dtype, device=device)

    def forward(self, x, mask, rngs, train=True):
        x = self.spatial_decompression(x)
        b, t, hw, _ = x.shape
        tmask = FactoredAttention.temporal_mask_arg(mask, b, t, hw)
        for layer in self.layers:
            x = layer(x, tmask)
        feats, rgb = self.patch_unembedding(x)
        # x + unet_output, fused (model.py:95-96); in bf16 feats is a view of a 16-channel-pitch map (UnembedFn)
        return self.unet(feats, residual=rgb, zero_padded=feats.stride(-2) != feats.shape[-1])


class VideoVAE(nn.Module):
    """train/model.py:101-136.  Returns (reconstruction, compressed_representation, selection, log_variance, mean)."""

    def __init__(self, height, width, channels, patch_size, encoder_depth, decoder_depth, mlp_dim, num_heads,
                 qkv_features, max_temporal_len, spatial_compression_rate, unembedding_upsample_rate, rngs,
                 dtype=torch.bfloat16, param_dtype=torch.float32, device=None):
        super().__init__()
        key = rngs.params()
        self.dtype = dtype
        self.encoder = Encoder(height, width, channels, patch_size, encoder_depth, mlp_dim, num_heads, qkv_features,
                               max_temporal_len, spatial_compression_rate, rngs, dtype, param_dtype, device=device)
        self.decoder = Decoder(height, width, channels, patch_size, decoder_depth, mlp_dim, num_heads, qkv_features,
                               max_temporal_len, spatial_compression_rate, unembedding_upsample_rate, rngs, dtype,
                               param_dtype, device=device)
        lat = channels * patch_size * patch_size // spatial_compression_rate
        fill = torch.randn(1, 1, 1, lat, generator=key, dtype=torch.float32) * 0.02
        self.fill_token = nn.Parameter(fill.to(_default_device(device)))

    def set_recompute(self, on=True):
        """Per-layer activation recompute (the reference's ``@nnx.remat`` on FactoredAttention, train/layers.py:209) for
        every transformer layer of the encoder and decoder; returns self."""
        for mod in self.modules():
            if isinstance(mod, FactoredAttention):
                mod.recompute = bool(on)
        return self

    def forward(self, x, mask, rngs, train=True, noise=None, gumbel_u=None):
        mean, log_variance, selection = self.encoder(x, mask, rngs, train=train, gumbel_u=gumbel_u)
        seed, offset = rngs.sampling() if (train and noise is None) else (0, 0)
        eps = noise.to(torch.float32).contiguous() if noise is not None else None
        c32, c_low = F_.ReparamGateFn.apply(mean, log_variance, selection, self.fill_token, eps, seed, offset, bool(train))
        reconstruction = self.decoder(c_low if c_low is not None else c32, mask, rngs, train=train)
        return reconstruction, c32, selection, log_variance, mean
