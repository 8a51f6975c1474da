"""Independent float64 numpy restatement of every primitive (TEST INFRASTRUCTURE ONLY).

Written from the published definitions (not from oracle/nn.py) so that the two
can be checked against each other in tests/test_oracle.py: the torch oracle is
checked against these (and so are the primitives of oracle/jaxshim, tests/test_jaxshim_cpu.py), since the
reference has no golden vectors and JAX is not installable here (the third-party
primitives are unpinned against real JAX; see oracle/__init__.py).
"""
import numpy as np


def layer_norm(x, scale=None, bias=None, eps=1e-6):
    x = np.asarray(x, np.float64)
    mu = x.mean(-1, keepdims=True)
    var = np.maximum((x * x).mean(-1, keepdims=True) - mu * mu, 0.0)
    y = (x - mu) / np.sqrt(var + eps)
    if scale is not None:
        y = y * scale
    if bias is not None:
        y = y + bias
    return y


def group_norm(x, groups, scale, bias, eps=1e-6):
    """x: [b, ..., c]; statistics per (sample, group) over every other axis."""
    x = np.asarray(x, np.float64)
    b, c = x.shape[0], x.shape[-1]
    out = np.empty_like(x)
    cg = c // groups
    for i in range(b):
        for g in range(groups):
            sl = x[i, ..., g * cg:(g + 1) * cg]
            mu = sl.mean()
            var = max((sl * sl).mean() - mu * mu, 0.0)
            out[i, ..., g * cg:(g + 1) * cg] = (sl - mu) / np.sqrt(var + eps)
    return out * scale + bias


def conv3d_same(x, k, bias=None):
    """x [b,t,h,w,ci], k [kt,kh,kw,ci,co]; cross-correlation, zero 'SAME' padding, stride 1."""
    x = np.asarray(x, np.float64)
    k = np.asarray(k, np.float64)
    kt, kh, kw, ci, co = k.shape
    b, t, h, w, _ = x.shape
    xp = np.pad(x, ((0, 0), (kt // 2, kt // 2), (kh // 2, kh // 2), (kw // 2, kw // 2), (0, 0)))
    y = np.zeros((b, t, h, w, co))
    for a in range(kt):
        for i in range(kh):
            for j in range(kw):
                y += xp[:, a:a + t, i:i + h, j:j + w, :] @ k[a, i, j]
    return y if bias is None else y + bias


def conv_transpose_122(x, k, bias=None):
    """jax.lax.conv_transpose(strides (1,2,2), 'SAME', transpose_kernel=False) written out:
    dilate the input by the stride, pad (k-1, k+s-2-(k-1)) = (1,1) per spatial axis, then
    cross-correlate with the UNflipped kernel."""
    x = np.asarray(x, np.float64)
    k = np.asarray(k, np.float64)
    b, t, h, w, ci = x.shape
    co = k.shape[-1]
    dil = np.zeros((b, t, 2 * h - 1, 2 * w - 1, ci))
    dil[:, :, ::2, ::2, :] = x
    dil = np.pad(dil, ((0, 0), (0, 0), (1, 1), (1, 1), (0, 0)))
    y = np.zeros((b, t, 2 * h, 2 * w, co))
    for i in range(2):
        for j in range(2):
            y += dil[:, :, i:i + 2 * h, j:j + 2 * w, :] @ k[0, i, j]
    return y if bias is None else y + bias


def max_pool_122(x):
    x = np.asarray(x, np.float64)
    b, t, h, w, c = x.shape
    return x.reshape(b, t, h // 2, 2, w // 2, 2, c).max(axis=(3, 5))


def rope_tables(head_dim, seq_len, base=10000.0):
    inv_freq = 1.0 / (base ** (np.arange(0, head_dim, 2, dtype=np.float64) / head_dim))
    freqs = np.outer(np.arange(seq_len, dtype=np.float64), inv_freq)
    emb = np.concatenate([freqs, freqs], axis=-1)
    return np.cos(emb), np.sin(emb)


def rope(x, base=10000.0):
    """x: [a, seq, heads, hd]; non-interleaved (half-split) rotation, positions 0..seq-1."""
    x = np.asarray(x, np.float64)
    hd = x.shape[-1]
    cos, sin = rope_tables(hd, x.shape[1], base)
    cos, sin = cos[None, :, None, :], sin[None, :, None, :]
    rot = np.concatenate([-x[..., hd // 2:], x[..., :hd // 2]], axis=-1)
    return x * cos + rot * sin


def attention(q, k, v, mask=None):
    """q,k,v [B,T,N,H]; mask broadcastable to [B,N,T,S], True = attend."""
    q, k, v = (np.asarray(a, np.float64) for a in (q, k, v))
    logits = np.einsum("btnh,bsnh->bnts", q, k) / np.sqrt(q.shape[-1])
    if mask is not None:
        logits = np.where(mask, logits, -0.7 * np.finfo(np.float32).max)
    logits = logits - logits.max(-1, keepdims=True)
    p = np.exp(logits)
    p = p / p.sum(-1, keepdims=True)
    return np.einsum("bnts,bsnh->btnh", p, v)


def softplus(x):
    return np.logaddexp(np.asarray(x, np.float64), 0.0)


def silu(x):
    x = np.asarray(x, np.float64)
    return x / (1.0 + np.exp(-x))
