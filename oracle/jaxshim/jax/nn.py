"""jax.nn subset."""
import math

import torch

from . import numpy as jnp
from ._core import asarray


def sigmoid(x):
    return torch.sigmoid(asarray(x))


def softplus(x):
    return jnp.logaddexp(x, 0.0)                            # jax.nn.softplus = logaddexp(x, 0)


def silu(x):
    x = asarray(x)
    return x * torch.sigmoid(x)


swish = silu


def relu(x):
    return torch.relu(asarray(x))


def softmax(x, axis=-1):
    return torch.softmax(asarray(x), dim=axis)


def dot_product_attention(query, key, value, bias=None, mask=None, *, scale=None, is_causal=False,
                          query_seq_lengths=None, key_value_seq_lengths=None, local_window_size=None, implementation=None):
    """jax.nn.dot_product_attention, XLA implementation (jax/_src/nn/functions.py::_dot_product_attention_core):
    q [B,T,N,H], k/v [B,S,K,H]; logits in >= fp32, * scale (default 1/sqrt(H)), + bias, masked entries replaced by
    -0.7 * finfo(logits.dtype).max, softmax in fp32, probabilities cast to key.dtype, contracted with value."""
    q, k, v = asarray(query), asarray(key), asarray(value)
    assert not is_causal and query_seq_lengths is None and key_value_seq_lengths is None and local_window_size is None
    B, T, N, H = q.shape
    assert k.shape[2] == N, "grouped-query attention is not part of the reference's path"
    scale = 1.0 / math.sqrt(H) if scale is None else scale
    logits_dtype = torch.promote_types(q.dtype, torch.float32)
    logits = torch.einsum("btnh,bsnh->bnts", q.to(logits_dtype), k.to(logits_dtype))
    logits = logits * scale
    if bias is not None:
        logits = (logits + asarray(bias)).to(logits_dtype)
    if mask is not None:
        m = asarray(mask)
        assert m.dtype == torch.bool and m.ndim == 4, "mask must be a 4-D boolean array broadcastable to [B,N,T,S]"
        large_negative = -0.7 * torch.finfo(logits_dtype).max
        logits = torch.where(m, logits, torch.full((), large_negative, dtype=logits_dtype))
    probs = torch.softmax(logits.to(torch.float32), dim=-1).to(k.dtype)
    return torch.einsum("bnts,bsnh->btnh", probs, v)
