"""B200-native drop-in for the reference's train/rl_model.py: the RL variant of the VAE.

Same kernels as ``model.py``; what changes is the head and the batch plumbing (rl_model.py:56-60, 119-147): the frame
gate is the probability ``sigmoid(logit)`` (b, t, 1), every sample is duplicated ('b ... -> (b 2) ...'), a Bernoulli
keep-mask is drawn per frame (Philox uniform < p), the decoder runs on the doubled batch, and the call returns
(reconstruction, compressed_representation, selection, selection_mask, log_variance, mean).
``noise`` / ``bernoulli_u`` inject the draws for parity tests.
"""
import torch
from torch import nn

from . import functional as F_
from . import ops
from .layers import FactoredAttention, _default_device
from .model import Decoder, Encoder as _Encoder


def repeat2(a):
    """einops 'b ... -> (b 2) ...' (each sample followed by its copy).  Not ``repeat_interleave``: with an integer
    count that op sizes its output through a host read, which a CUDA-graph capture forbids."""
    return a.unsqueeze(1).expand(a.shape[0], 2, *a.shape[1:]).reshape(2 * a.shape[0], *a.shape[1:])


class Encoder(_Encoder):
    """train/rl_model.py:14-60.  Returns (mean, log_variance, selection[b,t,1]) with selection a probability."""

    def forward(self, x, mask, rngs, train=True):
        x = self.patch_embedding(x)
        b, t, hw, _ = x.shape
        tmask = FactoredAttention.temporal_mask_arg(mask, b, t, hw)
        for layer in self.layers:
            x = layer(x, tmask)
        mean, log_variance, selection = F_.EncoderHeadFn.apply(
            x, self.dtype, "prob", float(self.gumbel_sigmoid.temperature), None, 0, 0,
            self.spatial_compression.kernel, self.spatial_compression.bias, self.variance_estimator.kernel,
            self.variance_estimator.bias, self.selection_layer1.kernel, self.selection_layer1.bias,
            self.selection_layer2.kernel, self.selection_layer2.bias)
        return mean, log_variance, selection.reshape(b, t, 1)


class VideoVAE(nn.Module):
    """train/rl_model.py:100-147."""

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

    def forward(self, x, mask, rngs, train=True, noise=None, bernoulli_u=None):
        mean, log_variance, selection = self.encoder(x, mask, rngs, train=train)
        b, t = selection.shape[:2]
        rep = repeat2                                                      # 'b ... -> (b 2) ...'
        selection = rep(selection).reshape(2 * b, t, 1, 1)
        mean2, lv2, mask2 = rep(mean), rep(log_variance), rep(mask)
        # one Gaussian draw per ORIGINAL sample (the reference samples the latent before duplicating it)
        seed, offset = rngs.sampling() if (train and noise is None) else (0, 0)
        if train:
            if noise is None:
                noise = ops.philox_fill_(torch.empty(mean.shape, dtype=torch.float32, device=mean.device), seed, offset,
                                         "normal")
            eps = rep(noise.to(torch.float32)).contiguous()
        else:
            eps = None
        if bernoulli_u is None:
            seed, offset = rngs.sampling()
            bernoulli_u = ops.philox_fill_(torch.empty(2 * b, t, dtype=torch.float32, device=mean.device), seed, offset,
                                           "uniform")
        selection_mask = (bernoulli_u.reshape(2 * b, t, 1, 1).to(torch.float32) < selection.detach()).to(torch.float32)
        c32, c_low = F_.ReparamGateFn.apply(mean2, lv2, selection_mask, self.fill_token, eps, 0, 0, bool(train))
        reconstruction = self.decoder(c_low if c_low is not None else c32, mask2, rngs, train=train)
        return reconstruction, c32, selection, selection_mask.to(c32.dtype), lv2, mean2
