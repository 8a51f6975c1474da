"""jax.lax subset, following the published lax algorithms (conv_general_dilated with explicit padding / lhs dilation,
conv_transpose = padding rule + lhs dilation + un-flipped correlation, reduce_window max with VALID padding)."""
import math

import torch
import torch.nn.functional as F

from ._core import Array, asarray


def dynamic_slice(operand, start_indices, slice_sizes):
    x = asarray(operand)
    idx = []
    for dim, (s, n) in enumerate(zip(start_indices, slice_sizes)):
        s = max(0, min(int(s), x.shape[dim] - n))          # lax clamps the start so the slice stays in bounds
        idx.append(slice(s, s + n))
    return x[tuple(idx)]


def rsqrt(x):
    return torch.rsqrt(asarray(x))


def stop_gradient(x):
    return asarray(x).detach()


def _same_pads(in_size, k, stride):
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return total // 2, total - total // 2


def _dilate(x, strides):
    """lhs dilation of a channels-last [N, *spatial, C] tensor: stride - 1 zeros between neighbouring elements."""
    if all(s == 1 for s in strides):
        return x
    sp = x.shape[1:-1]
    out = x.new_zeros((x.shape[0],) + tuple((n - 1) * s + 1 for n, s in zip(sp, strides)) + (x.shape[-1],))
    out[(slice(None),) + tuple(slice(None, None, s) for s in strides) + (slice(None),)] = x
    return out


def conv_general_dilated_nhwc(lhs, rhs, window_strides, padding, lhs_dilation=None):
    """lhs [N, *spatial, Cin] (x) rhs [*k, Cin, Cout] -> [N, *spatial', Cout]: cross-correlation, no kernel flip."""
    nd = lhs.ndim - 2
    ks = rhs.shape[:nd]
    if lhs_dilation is not None:
        lhs = _dilate(lhs, lhs_dilation)
    if isinstance(padding, str):
        if padding.upper() == "SAME":
            padding = [_same_pads(lhs.shape[1 + i], ks[i], window_strides[i]) for i in range(nd)]
        elif padding.upper() == "VALID":
            padding = [(0, 0)] * nd
        else:
            raise NotImplementedError(padding)
    flat = []
    for lo, hi in reversed(list(padding)):                  # F.pad wants the last dimension first
        flat += [lo, hi]
    perm_in = (0, nd + 1) + tuple(range(1, nd + 1))         # channels-last -> channels-first
    x = F.pad(lhs.permute(*perm_in), flat)
    w = rhs.permute(nd + 1, nd, *range(nd))                 # [*k, I, O] -> [O, I, *k]
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    y = conv(x, w, stride=tuple(window_strides))
    return y.permute(0, *range(2, nd + 2), 1)


def _conv_transpose_padding(k, s, padding):
    """jax/_src/lax/convolution.py::_conv_transpose_padding."""
    if padding == "SAME":
        pad_len = k + s - 2
        pad_a = k - 1 if s > k - 1 else int(math.ceil(pad_len / 2))
    elif padding == "VALID":
        pad_len = k + s - 2 + max(k - s, 0)
        pad_a = k - 1
    else:
        raise ValueError(padding)
    return pad_a, pad_len - pad_a


def conv_transpose_nhwc(lhs, rhs, strides, padding, transpose_kernel=False):
    """lax.conv_transpose with channels-last dimension numbers (what flax's ConvTranspose calls)."""
    nd = lhs.ndim - 2
    ks = rhs.shape[:nd]
    if isinstance(padding, str):
        padding = [_conv_transpose_padding(k, s, padding.upper()) for k, s in zip(ks, strides)]
    if transpose_kernel:
        rhs = torch.flip(rhs, dims=tuple(range(nd))).transpose(nd, nd + 1)
    return conv_general_dilated_nhwc(lhs, rhs, (1,) * nd, padding, lhs_dilation=tuple(strides))


def reduce_window_max_nhwc(x, window, strides):
    """flax pool(..., -inf, lax.max, window, strides, 'VALID') on [N, *spatial, C]."""
    nd = x.ndim - 2
    perm_in = (0, nd + 1) + tuple(range(1, nd + 1))
    pool = {1: F.max_pool1d, 2: F.max_pool2d, 3: F.max_pool3d}[nd]
    y = pool(x.permute(*perm_in), kernel_size=tuple(window), stride=tuple(strides))
    return y.permute(0, *range(2, nd + 2), 1)
