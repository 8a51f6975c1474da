"""Oracle restatement of train/layers.py (CPU torch, TEST INFRASTRUCTURE ONLY).

Class names, constructor arguments, call signatures and parameter names follow
the reference one for one; each class cites the lines it restates.
"""
import math

import torch
from torch import nn

from . import nn as onn


class PatchEmbedding(nn.Module):
    """train/layers.py:8-27 -- rearrange -> cast -> LayerNorm -> Linear."""

    def __init__(self, height, width, channels, patch_size, rngs, dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.patch_size = patch_size
        self.dtype = dtype
        d = patch_size * patch_size * channels
        self.linear = onn.Linear(d, d, rngs, dtype, param_dtype)
        self.norm = onn.LayerNorm(d, rngs, dtype, param_dtype)

    def forward(self, x):
        b, t, H, W, c = x.shape
        p = self.patch_size
        # "b t (h p1) (w p2) c -> b t (h w) (p1 p2 c)"
        x = x.reshape(b, t, H // p, p, W // p, p, c).permute(0, 1, 2, 4, 3, 5, 6).reshape(b, t, (H // p) * (W // p), p * p * c)
        x = x.to(self.dtype)
        return self.linear(self.norm(x))


class PatchUnEmbedding(nn.Module):
    """train/layers.py:29-55 -- Linear, Linear(x upsample), pixel shuffle, Linear(C*u -> C)."""

    def __init__(self, height, width, channels, patch_size, upsample_rate, rngs,
                 dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.patch_size, self.height, self.width, self.upsample_rate = patch_size, height, width, upsample_rate
        d = patch_size * patch_size * channels
        self.upsample = onn.Linear(d, d * upsample_rate, rngs, dtype, param_dtype)
        self.downsample = onn.Linear(channels * upsample_rate, channels, rngs, dtype, param_dtype)
        self.linear = onn.Linear(d, d, rngs, dtype, param_dtype)

    def forward(self, x):
        b, t = x.shape[:2]
        p, u = self.patch_size, self.upsample_rate
        h, w = self.height // p, self.width // p
        x = self.upsample(self.linear(x))
        cu = x.shape[-1] // (p * p)
        # "b t (h w) (p1 p2 c u) -> b t (h p1) (w p2) (c u)"
        feats = x.reshape(b, t, h, w, p, p, cu).permute(0, 1, 2, 4, 3, 5, 6).reshape(b, t, h * p, w * p, cu)
        return feats, self.downsample(feats)


def rotate_half(x):
    """train/layers.py:80-83."""
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


class RotaryEmbedding(nn.Module):
    """train/layers.py:85-129 (tables are non-trainable state, positions 0..seq-1)."""

    def __init__(self, head_dim, max_len=8192, alpha=1.0, base=10000.0):
        super().__init__()
        self.head_dim, self.max_len = head_dim, max_len
        ntk_base = base * (alpha ** (head_dim / (head_dim - 2))) if head_dim != 2 else base
        inv_freq = 1.0 / (ntk_base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
        t = torch.arange(max_len, dtype=torch.float32)
        freqs = torch.einsum("i,j->ij", t, inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        self.register_buffer("cos_cached", torch.cos(emb)[None, :, None, :], persistent=False)
        self.register_buffer("sin_cached", torch.sin(emb)[None, :, None, :], persistent=False)

    def rotate_queries_and_keys(self, q, k):
        seq_len = q.shape[1]
        cos = self.cos_cached[:, :seq_len].to(q.dtype)
        sin = self.sin_cached[:, :seq_len].to(q.dtype)
        return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


class Attention(nn.Module):
    """train/layers.py:131-171 (use_qk_norm is ignored by the reference: QK-norm always on)."""

    def __init__(self, in_features, num_heads, qkv_features, max_len, use_qk_norm, rngs,
                 dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.num_heads = num_heads
        head_dim = qkv_features // num_heads
        self.qkv_projection = onn.Linear(in_features, qkv_features * 3, rngs, dtype, param_dtype)
        self.out_projection = onn.Linear(qkv_features, in_features, rngs, dtype, param_dtype, init_scale=1e-2)
        self.input_norm = onn.LayerNorm(in_features, rngs, dtype, param_dtype)
        self.ROPE = RotaryEmbedding(head_dim=head_dim, max_len=max_len)
        self.use_qk_norm = use_qk_norm
        self.q_norm = onn.LayerNorm(head_dim, rngs, dtype, param_dtype, use_bias=False)
        self.k_norm = onn.LayerNorm(head_dim, rngs, dtype, param_dtype, use_bias=False)

    def forward(self, x, mask=None):
        a, s, _ = x.shape
        x = self.input_norm(x)
        q, k, v = torch.chunk(self.qkv_projection(x), 3, dim=-1)
        q = q.reshape(a, s, self.num_heads, -1)
        k = k.reshape(a, s, self.num_heads, -1)
        v = v.reshape(a, s, self.num_heads, -1)
        q, k = self.q_norm(q), self.k_norm(k)
        q, k = self.ROPE.rotate_queries_and_keys(q, k)
        o = onn.dot_product_attention(q, k, v, mask=mask)
        return self.out_projection(o.reshape(a, s, -1))


class MLP(nn.Module):
    """train/layers.py:174-196."""

    def __init__(self, in_features, mlp_dim, rngs, dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.norm = onn.LayerNorm(in_features, rngs, dtype, param_dtype)
        self.linear1 = onn.Linear(in_features, mlp_dim, rngs, dtype, param_dtype)
        self.linear2 = onn.Linear(mlp_dim, in_features, rngs, dtype, param_dtype, init_scale=1e-2)

    def forward(self, x):
        return self.linear2(onn.silu(self.linear1(self.norm(x))))


class FactoredAttention(nn.Module):
    """train/layers.py:198-224; also accepts the (b,1,1,t) mask of claude_distributed/layers.py:213-214."""

    def __init__(self, mlp_dim, in_features, num_heads, qkv_features, max_temporal_len, max_spatial_len, rngs,
                 dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.SpatialAttention = Attention(in_features, num_heads, qkv_features, max_spatial_len, True, rngs, dtype, param_dtype)
        self.SpatialMLP = MLP(in_features, mlp_dim, rngs, dtype, param_dtype)
        self.TemporalAttention = Attention(in_features, num_heads, qkv_features, max_temporal_len, False, rngs, dtype, param_dtype)
        self.TemporalMLP = MLP(in_features, mlp_dim, rngs, dtype, param_dtype)

    def forward(self, x, temporal_mask):
        b, t, hw, c = x.shape
        if temporal_mask is not None and temporal_mask.shape[0] == b and hw != 1:
            temporal_mask = temporal_mask.repeat_interleave(hw, dim=0)  # (b,1,1,t) -> ((b hw),1,1,t)
        tx = x.permute(0, 2, 1, 3).reshape(b * hw, t, c)
        tx = tx + self.TemporalAttention(tx, mask=temporal_mask)
        tx = tx + self.TemporalMLP(tx)
        x = tx.reshape(b, hw, t, c).permute(0, 2, 1, 3)
        sx = x.reshape(b * t, hw, c)
        sx = sx + self.SpatialAttention(sx)
        sx = sx + self.SpatialMLP(sx)
        return sx.reshape(b, t, hw, c)


class _RoundSTE(torch.autograd.Function):
    """train/layers.py:226-236 -- round (half to even) forward, identity backward."""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


def round_ste(logits):
    return _RoundSTE.apply(logits)


class GumbelSigmoidSTE(nn.Module):
    """train/layers.py:238-252.  ``u`` optionally injects the uniform draw (parity tests)."""

    def __init__(self, temperature: float = 1.0):
        super().__init__()
        self.temperature = temperature

    def forward(self, logits, rngs, train=True, u=None):
        if train:
            if u is None:
                u = torch.rand(logits.shape, generator=rngs.sampling(), dtype=torch.float32)
            eps = 1e-20
            u = torch.clamp(u.to(torch.float32), eps, 1.0 - eps)
            noise = torch.log(u / (1 - u))
            wide = torch.promote_types(logits.dtype, torch.float32)  # jnp promotion: bf16 + f32 -> f32
            return round_ste(torch.sigmoid((logits.to(wide) + noise.to(wide)) / self.temperature))
        return torch.round(torch.sigmoid(logits / self.temperature))


def sqrt_head_scale(head_dim):
    return 1.0 / math.sqrt(head_dim)
