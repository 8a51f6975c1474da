"""B200-native drop-in for the reference's train/layers.py.

Same class names, constructor arguments, call signatures and Flax parameter names (kernel / bias / scale) as the
reference; the arithmetic runs in hand-written sm_100a kernels through libvvae (no CPU / PyTorch-op path).
Activations stay in the canonical token layout [b, t, hw, c] for the whole transformer stack: temporal attention reads
its sequences strided (stride hw) instead of materialising the reference's "(b hw) t c" transposes
(train/layers.py:212,217,219,224).
"""
import math

import torch
from torch import nn

from . import functional as F_
from . import ops
from .functional import AttnCfg
from .ops import AttnGeom, AttnMask

_TRUNC = 0.87962566103423978


def _default_device(device):
    if device is not None:
        return torch.device(device)
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _variance_scaling(shape, fan_in, scale, gen):
    """flax variance_scaling(scale, 'fan_in', 'truncated_normal') (lecun_normal when scale == 1)."""
    std = math.sqrt(scale / fan_in) / _TRUNC
    t = torch.empty(shape, dtype=torch.float32)
    nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    return t


class Linear(nn.Module):
    """nnx.Linear parameter holder: kernel (in,out), bias (out)."""

    def __init__(self, in_features, out_features, rngs, dtype=torch.bfloat16, param_dtype=torch.float32,
                 init_scale=1.0, device=None):
        super().__init__()
        assert param_dtype == torch.float32, "parameters are kept in float32"
        dev = _default_device(device)
        self.dtype = dtype
        self.kernel = nn.Parameter(_variance_scaling((in_features, out_features), in_features, init_scale,
                                                     rngs.params()).to(dev))
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=torch.float32, device=dev))

    def forward(self, x):
        return F_.LinearFn.apply(x, self.kernel, self.bias, self.dtype)


class LayerNorm(nn.Module):
    """nnx.LayerNorm(eps=1e-6) parameter holder: scale, bias."""

    def __init__(self, features, rngs=None, dtype=torch.bfloat16, param_dtype=torch.float32, use_bias=True, device=None):
        super().__init__()
        dev = _default_device(device)
        self.dtype = dtype
        self.scale = nn.Parameter(torch.ones(features, dtype=torch.float32, device=dev))
        self.bias = nn.Parameter(torch.zeros(features, dtype=torch.float32, device=dev)) if use_bias else None

    def forward(self, x):
        return F_.LayerNormFn.apply(x, self.scale, self.bias, self.dtype)


class PatchEmbedding(nn.Module):
    """train/layers.py:8-27."""

    def __init__(self, height, width, channels, patch_size, rngs, dtype=torch.bfloat16, param_dtype=torch.float32,
                 device=None):
        super().__init__()
        self.patch_size, self.dtype = patch_size, dtype
        d = patch_size * patch_size * channels
        self.linear = Linear(d, d, rngs, dtype, param_dtype, device=device)
        self.norm = LayerNorm(d, rngs, dtype, param_dtype, device=device)

    def forward(self, x):
        return F_.PatchEmbedFn.apply(x, self.patch_size, self.dtype, self.norm.scale, self.norm.bias,
                                     self.linear.kernel, self.linear.bias)


class PatchUnEmbedding(nn.Module):
    """train/layers.py:29-55; returns (convolutional_upsampled_features, x)."""

    def __init__(self, height, width, channels, patch_size, upsample_rate, rngs, dtype=torch.bfloat16,
                 param_dtype=torch.float32, device=None):
        super().__init__()
        self.patch_size, self.height, self.width, self.upsample_rate = patch_size, height, width, upsample_rate
        self.channels, self.dtype = channels, dtype
        d = patch_size * patch_size * channels
        self.upsample = Linear(d, d * upsample_rate, rngs, dtype, param_dtype, device=device)
        self.downsample = Linear(channels * upsample_rate, channels, rngs, dtype, param_dtype, device=device)
        self.linear = Linear(d, d, rngs, dtype, param_dtype, device=device)

    def forward(self, x):
        b, t = x.shape[:2]
        geo = (b, t, self.height, self.width, self.patch_size, self.channels * self.upsample_rate)
        if x.dtype != self.dtype:
            x = ops.cast(x.contiguous(), self.dtype)
        return F_.UnembedFn.apply(x, geo, self.dtype, self.linear.kernel, self.linear.bias, self.upsample.kernel,
                                  self.upsample.bias, self.downsample.kernel, self.downsample.bias)


def rotate_half(x):
    """train/layers.py:80-83 (kept for API parity; the kernels fuse it with QK-norm)."""
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


class RotaryEmbedding(nn.Module):
    """train/layers.py:85-129: tables cos/sin [max_len, head_dim] (fp32 state, not trained)."""

    def __init__(self, head_dim, max_len=8192, alpha=1.0, base=10000.0, device=None):
        super().__init__()
        self.head_dim, self.max_len = head_dim, max_len
        ntk_base = base * (alpha ** (head_dim / (head_dim - 2))) if head_dim != 2 else base
        inv_freq = 1.0 / (ntk_base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
        freqs = torch.einsum("i,j->ij", torch.arange(max_len, dtype=torch.float32), inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        dev = _default_device(device)
        self.register_buffer("cos_cached", torch.cos(emb).contiguous().to(dev), persistent=False)
        self.register_buffer("sin_cached", torch.sin(emb).contiguous().to(dev), persistent=False)
        self._typed = {}

    def tables(self, dtype):
        """cos/sin in the compute dtype (the reference casts them to q's dtype, train/layers.py:124-127)."""
        if dtype == torch.float32:
            return self.cos_cached, self.sin_cached
        ent = self._typed.get(dtype)
        if ent is None or ent[0].device != self.cos_cached.device:
            from . import ops
            ent = (ops.cast(self.cos_cached, dtype), ops.cast(self.sin_cached, dtype))
            self._typed[dtype] = ent
        return ent


def _mask_rows(mask, n_rows, L):
    """Reduce a reference-style bool mask [a,1,1,L] (or [a,L]) to uint8 [a, L]."""
    m = mask.reshape(mask.shape[0], -1)
    if m.shape[1] != L:
        raise ValueError(f"mask last dim {m.shape[1]} != sequence length {L}")
    return m.to(torch.uint8).contiguous()


class Attention(nn.Module):
    """train/layers.py:131-171.  ``__call__(x[a,seq,dim], mask[a,1,1,seq] | None)``; QK-norm is always applied."""

    def __init__(self, in_features, num_heads, qkv_features, max_len, use_qk_norm, rngs, dtype=torch.bfloat16,
                 param_dtype=torch.float32, device=None):
        super().__init__()
        self.num_heads, self.dtype = num_heads, dtype
        self.head_dim = qkv_features // num_heads
        self.qkv_projection = Linear(in_features, qkv_features * 3, rngs, dtype, param_dtype, device=device)
        self.out_projection = Linear(qkv_features, in_features, rngs, dtype, param_dtype, init_scale=1e-2, device=device)
        self.input_norm = LayerNorm(in_features, rngs, dtype, param_dtype, device=device)
        self.ROPE = RotaryEmbedding(head_dim=self.head_dim, max_len=max_len, device=device)
        self.use_qk_norm = use_qk_norm
        self.q_norm = LayerNorm(self.head_dim, rngs, dtype, param_dtype, use_bias=False, device=device)
        self.k_norm = LayerNorm(self.head_dim, rngs, dtype, param_dtype, use_bias=False, device=device)

    def _run(self, x, cfg):
        return F_.AttnBlockFn.apply(x, cfg, self.input_norm.scale, self.input_norm.bias, self.qkv_projection.kernel,
                                    self.qkv_projection.bias, self.q_norm.scale, self.k_norm.scale,
                                    self.out_projection.kernel, self.out_projection.bias, *self.ROPE.tables(self.dtype))

    def forward(self, x, mask=None):
        a, seq, _ = x.shape
        if seq > self.ROPE.max_len:
            raise ValueError(f"sequence length {seq} exceeds RoPE table {self.ROPE.max_len}")
        geom = AttnGeom(a, 1, seq, seq, 0, 1)
        am = None
        if mask is not None:
            if mask.dim() == 4 and (mask.shape[1] != 1 or mask.shape[2] != 1):
                # general [a, h|1, q|1, k] mask (train/attention_mask_tests.py)
                m = mask.to(torch.uint8).contiguous()
                st = [m.stride(i) if m.shape[i] != 1 else 0 for i in range(4)]
                am = AttnMask(m, 1, st[0], st[1], st[2], st[3])
            else:
                am = AttnMask(_mask_rows(mask, a, seq), 1, seq, 0, 0, 1)
        cfg = AttnCfg(geom, self.num_heads, self.head_dim, 1, seq, am, False, self.dtype)
        return self._run(x, cfg)


class MLP(nn.Module):
    """train/layers.py:174-196."""

    def __init__(self, in_features, mlp_dim, rngs, dtype=torch.bfloat16, param_dtype=torch.float32, device=None):
        super().__init__()
        self.dtype = dtype
        self.norm = LayerNorm(in_features, rngs, dtype, param_dtype, device=device)
        self.linear1 = Linear(in_features, mlp_dim, rngs, dtype, param_dtype, device=device)
        self.linear2 = Linear(mlp_dim, in_features, rngs, dtype, param_dtype, init_scale=1e-2, device=device)

    def _run(self, x, residual):
        return F_.MlpBlockFn.apply(x, residual, self.dtype, self.norm.scale, self.norm.bias, self.linear1.kernel,
                                   self.linear1.bias, self.linear2.kernel, self.linear2.bias)

    def forward(self, x):
        return self._run(x, False)


class FactoredAttention(nn.Module):
    """train/layers.py:198-224.  ``__call__(x[b,t,hw,c], temporal_mask)`` with temporal_mask ((b hw),1,1,t) as passed
    by train/, (b,1,1,t) as passed by claude_distributed/ (layers.py:213-214), (b,t), or None."""

    def __init__(self, mlp_dim, in_features, num_heads, qkv_features, max_temporal_len, max_spatial_len, rngs,
                 dtype=torch.bfloat16, param_dtype=torch.float32, device=None, recompute=False):
        super().__init__()
        self.dtype = dtype
        # the reference's @nnx.remat (train/layers.py:209) as a switch: keep only the layer input for backward and re-run
        # the forward kernels there (saves ~1.1 GB per layer at 16x256x256 x 8 clips for one extra forward)
        self.recompute = bool(recompute)
        self.SpatialAttention = Attention(in_features, num_heads, qkv_features, max_spatial_len, True, rngs, dtype,
                                          param_dtype, device=device)
        self.SpatialMLP = MLP(in_features, mlp_dim, rngs, dtype, param_dtype, device=device)
        self.TemporalAttention = Attention(in_features, num_heads, qkv_features, max_temporal_len, False, rngs, dtype,
                                           param_dtype, device=device)
        self.TemporalMLP = MLP(in_features, mlp_dim, rngs, dtype, param_dtype, device=device)

    @staticmethod
    def temporal_mask_arg(mask, b, t, hw):
        """Normalise the reference's mask conventions to an AttnMask over sequences (b, hw) without expanding it."""
        if mask is None:
            return None
        m = mask.reshape(mask.shape[0], -1)
        if m.shape[1] != t:
            raise ValueError(f"temporal mask has {m.shape[1]} frames, input has {t}")
        m = m.to(torch.uint8).contiguous()
        if m.shape[0] == b * hw and hw != 1:
            return AttnMask(m, 1, t, 0, 0, 1)          # one row per (b, hw) sequence
        if m.shape[0] == b:
            return AttnMask(m, hw, t, 0, 0, 1)         # one row per clip, shared by its hw sequences
        raise ValueError(f"temporal mask batch {m.shape[0]} matches neither b={b} nor b*hw={b * hw}")

    def forward(self, x, temporal_mask):
        b, t, hw, c = x.shape
        ta, sa = self.TemporalAttention, self.SpatialAttention
        if t > ta.ROPE.max_len or hw > sa.ROPE.max_len:
            raise ValueError("sequence longer than the RoPE tables (max_temporal_len / max_spatial_len)")
        tmask = temporal_mask if isinstance(temporal_mask, AttnMask) else self.temporal_mask_arg(temporal_mask, b, t, hw)
        # temporal: sequences (b, hw), positions t, tokens hw apart; RoPE position = frame index
        tcfg = AttnCfg(AttnGeom(b, hw, t, t * hw, 1, hw), ta.num_heads, ta.head_dim, hw, t, tmask, True, self.dtype)
        # spatial: sequences (b t), positions hw, contiguous; RoPE position = flattened patch index
        scfg = AttnCfg(AttnGeom(b * t, 1, hw, hw, 0, 1), sa.num_heads, sa.head_dim, 1, hw, None, True, self.dtype)

        def run(h):
            h = ta._run(h, tcfg)
            h = self.TemporalMLP._run(h, True)
            h = sa._run(h, scfg)
            return self.SpatialMLP._run(h, True)

        if self.recompute and torch.is_grad_enabled():
            return F_.RecomputeFn.apply(x, run, *[p for p in self.parameters() if p.requires_grad])
        return run(x)


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


def round_ste(logits):
    """train/layers.py:226-236 (a [b,t]-sized scalar op; kept in torch for API parity)."""
    return _RoundSTE.apply(logits)


class GumbelSigmoidSTE(nn.Module):
    """train/layers.py:238-252.  Stand-alone version for API parity; inside Encoder the gate is fused into
    vvae_selection_fwd.  ``rngs`` is a video_vae_b200.Rngs; ``u`` optionally injects the uniform draw."""

    def __init__(self, temperature: float = 1.0):
        super().__init__()
        self.temperature = temperature

    def forward(self, logits, rngs, train=True, u=None):
        if train:
            if u is None:
                seed, offset = rngs.sampling()
                g = torch.Generator(device=logits.device).manual_seed((seed ^ offset) & 0x7FFFFFFFFFFFFFFF)
                u = torch.rand(logits.shape, generator=g, device=logits.device, dtype=torch.float32)
            u = torch.clamp(u.float(), 1e-20, 1.0 - 1e-20)
            noise = torch.log(u / (1 - u))
            return round_ste(torch.sigmoid((logits.float() + noise) / self.temperature))
        return torch.round(torch.sigmoid(logits / self.temperature))
