"""Oracle primitives: the Flax-NNX / JAX defaults the reference relies on (CPU torch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every class keeps the Flax
parameter names (kernel / bias / scale) and layouts (Linear kernel (in,out), conv
kernel (kt,kh,kw,Cin,Cout)) so state_dicts map 1:1 onto the reference's param
tree.

Defaults restated here (flax 0.12.4 / jax 0.9.0.1, un-vendored):
  * Linear:    y = x @ K + b, inputs and params cast to ``dtype`` first.
  * LayerNorm: eps 1e-6, statistics in >=fp32 with the "fast variance"
               max(0, E[x^2] - E[x]^2); normalise in fp32, cast to ``dtype``.
  * GroupNorm: same rule, reduction over every non-batch axis inside a group.
  * Conv:      NDHWC cross-correlation, 'SAME' zero padding, stride 1.
  * ConvTranspose k = s = (1,2,2), 'SAME', transpose_kernel=False:
               out[2i + a] = x[i] . K[1 - a] per spatial axis (taps flipped
               relative to torch.nn.ConvTranspose3d).
  * max_pool:  VALID, first maximum wins on ties.
  * dot_product_attention: logits fp32, scale 1/sqrt(H), masked logits replaced
               by -0.7 * finfo(f32).max, softmax fp32, probabilities cast to the
               key dtype before the PV contraction.
  * lecun_normal = truncated normal (+-2 sigma), sigma = sqrt(scale / fan_in) / 0.87962566.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

_TRUNC_STD_CORRECTION = 0.87962566103423978


def variance_scaling_(t: torch.Tensor, fan_in: int, scale: float = 1.0, generator=None):
    """flax.nnx.initializers.variance_scaling(scale, 'fan_in', 'truncated_normal')."""
    std = math.sqrt(scale / fan_in) / _TRUNC_STD_CORRECTION
    with torch.no_grad():
        nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0 * std, b=2.0 * std, generator=generator)
    return t


def _stat_dtype(x):
    return torch.promote_types(torch.float32, x.dtype)


class Linear(nn.Module):
    """nnx.Linear (call sites train/layers.py:15,38-43,138-150,178-189; train/model.py:27-34,72)."""

    def __init__(self, in_features, out_features, rngs, dtype=torch.float32, param_dtype=torch.float32,
                 init_scale: float = 1.0, use_bias: bool = True):
        super().__init__()
        self.dtype = dtype
        self.kernel = nn.Parameter(torch.empty(in_features, out_features, dtype=param_dtype))
        variance_scaling_(self.kernel, in_features, init_scale, rngs.params())
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=param_dtype)) if use_bias else None

    def forward(self, x):
        y = x.to(self.dtype) @ self.kernel.to(self.dtype)
        if self.bias is not None:
            y = y + self.bias.to(self.dtype)
        return y


def layer_norm(x, scale, bias, dtype, eps=1e-6):
    xs = x.to(_stat_dtype(x))
    mean = xs.mean(-1, keepdim=True)
    mean2 = (xs * xs).mean(-1, keepdim=True)
    var = torch.clamp(mean2 - mean * mean, min=0.0)
    y = (xs - mean) * torch.rsqrt(var + eps)
    if scale is not None:
        y = y * scale.to(y.dtype)
    if bias is not None:
        y = y + bias.to(y.dtype)
    return y.to(dtype)


class LayerNorm(nn.Module):
    """nnx.LayerNorm(eps=1e-6) (train/layers.py:17,153,155-156,178)."""

    def __init__(self, features, rngs=None, dtype=torch.float32, param_dtype=torch.float32, use_bias=True):
        super().__init__()
        self.dtype = dtype
        self.scale = nn.Parameter(torch.ones(features, dtype=param_dtype))
        self.bias = nn.Parameter(torch.zeros(features, dtype=param_dtype)) if use_bias else None

    def forward(self, x):
        return layer_norm(x, self.scale, self.bias, self.dtype)


class GroupNorm(nn.Module):
    """nnx.GroupNorm(num_groups, eps=1e-6) on channels-last input (train/unet.py:22-23)."""

    def __init__(self, num_groups, num_features, rngs=None, dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.dtype = dtype
        self.num_groups = num_groups
        self.scale = nn.Parameter(torch.ones(num_features, dtype=param_dtype))
        self.bias = nn.Parameter(torch.zeros(num_features, dtype=param_dtype))

    def forward(self, x):
        b, c = x.shape[0], x.shape[-1]
        g = self.num_groups
        xs = x.to(_stat_dtype(x)).reshape(b, -1, g, c // g)
        mean = xs.mean(dim=(1, 3), keepdim=True)
        mean2 = (xs * xs).mean(dim=(1, 3), keepdim=True)
        var = torch.clamp(mean2 - mean * mean, min=0.0)
        y = ((xs - mean) * torch.rsqrt(var + 1e-6)).reshape(x.shape)
        y = y * self.scale.to(y.dtype) + self.bias.to(y.dtype)
        return y.to(self.dtype)


class Conv(nn.Module):
    """nnx.Conv, NDHWC, 'SAME' (train/unet.py:13-21,111-113,144-153)."""

    def __init__(self, in_features, out_features, kernel_size, rngs, dtype=torch.float32,
                 param_dtype=torch.float32, zero_init=False):
        super().__init__()
        self.dtype = dtype
        kt, kh, kw = kernel_size
        self.kernel = nn.Parameter(torch.zeros(kt, kh, kw, in_features, out_features, dtype=param_dtype))
        if not zero_init:
            variance_scaling_(self.kernel, kt * kh * kw * in_features, 1.0, rngs.params())
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=param_dtype))

    def forward(self, x):
        kt, kh, kw = self.kernel.shape[:3]
        w = self.kernel.to(self.dtype).permute(4, 3, 0, 1, 2)
        y = F.conv3d(x.to(self.dtype).permute(0, 4, 1, 2, 3), w, None, padding=(kt // 2, kh // 2, kw // 2))
        return y.permute(0, 2, 3, 4, 1) + self.bias.to(self.dtype)


class ConvTranspose122(nn.Module):
    """nnx.ConvTranspose(kernel (1,2,2), strides (1,2,2)) (train/unet.py:61-69)."""

    def __init__(self, in_features, out_features, rngs, dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.dtype = dtype
        self.kernel = nn.Parameter(torch.empty(1, 2, 2, in_features, out_features, dtype=param_dtype))
        variance_scaling_(self.kernel, 4 * in_features, 1.0, rngs.params())
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=param_dtype))

    def forward(self, x):
        b, t, h, w, _ = x.shape
        k = self.kernel.to(self.dtype)[0].flip(0, 1)  # out[2i+a, 2j+c] uses K[1-a, 1-c]
        y = torch.einsum("bthwi,acio->bthawco", x.to(self.dtype), k)
        y = y.reshape(b, t, 2 * h, 2 * w, -1)
        return y + self.bias.to(self.dtype)


def max_pool_122(x):
    """nnx.max_pool(window (1,2,2), strides (1,2,2)) (train/unet.py:50)."""
    y = F.max_pool3d(x.permute(0, 4, 1, 2, 3), kernel_size=(1, 2, 2), stride=(1, 2, 2))
    return y.permute(0, 2, 3, 4, 1)


def silu(x):
    return x * torch.sigmoid(x)


def softplus(x):
    """jax.nn.softplus = logaddexp(x, 0) (train/model.py:54)."""
    return torch.logaddexp(x, torch.zeros((), dtype=x.dtype))


def dot_product_attention(q, k, v, mask=None):
    """jax.nn.dot_product_attention, layout [B, T, N, H] (train/layers.py:168)."""
    b, t, n, h = q.shape
    ld = _stat_dtype(q)
    logits = torch.einsum("btnh,bsnh->bnts", q.to(ld), k.to(ld)) * (1.0 / math.sqrt(h))
    if mask is not None:
        big_neg = -0.7 * torch.finfo(torch.float32).max
        logits = torch.where(mask.to(torch.bool), logits, torch.full((), big_neg, dtype=ld))
    probs = torch.softmax(logits, dim=-1).to(k.dtype)
    out = torch.einsum("bnts,bsnh->btnh", probs.to(ld), v.to(ld))
    return out.to(v.dtype)
