"""B200-native drop-in for the reference's train/unet.py (3-D conv U-Net refiner of the decoder).

Same classes / constructor arguments / parameter names as the reference.  ``UNet.forward`` runs as ONE autograd
Function whose backward is written out by hand: skip tensors are produced directly inside the concat buffers of the
decoder path (no torch.cat copy), the skip gradient is folded into the max-pool backward, the residual
``x + unet_output`` of train/model.py:96 is fused into final_conv's epilogue.
"""
import torch
from torch import nn
from torch.autograd import Function

from . import functional as F_
from . import ops
from ._ffi import require_device
from .layers import _default_device, _variance_scaling


class Conv(nn.Module):
    """nnx.Conv parameter holder: kernel (kt,kh,kw,Cin,Cout), bias (Cout); NDHWC, 'SAME', stride 1."""

    def __init__(self, in_features, out_features, kernel_size, rngs, dtype=torch.bfloat16, param_dtype=torch.float32,
                 zero_init=False, device=None):
        super().__init__()
        dev = _default_device(device)
        self.dtype = dtype
        kt, kh, kw = kernel_size
        shape = (kt, kh, kw, in_features, out_features)
        k = torch.zeros(shape, dtype=torch.float32) if zero_init else _variance_scaling(
            shape, kt * kh * kw * in_features, 1.0, rngs.params())
        self.kernel = nn.Parameter(k.to(dev))
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=torch.float32, device=dev))

    @property
    def ks(self):
        return tuple(self.kernel.shape[:3])

    def forward(self, x):
        return ConvFn.apply(x, self.kernel, self.bias, self.dtype)


class ConvTranspose(nn.Module):
    """nnx.ConvTranspose(kernel (1,2,2), strides (1,2,2)) parameter holder."""

    def __init__(self, in_features, out_features, rngs, dtype=torch.bfloat16, param_dtype=torch.float32, device=None):
        super().__init__()
        dev = _default_device(device)
        self.dtype = dtype
        self.kernel = nn.Parameter(_variance_scaling((1, 2, 2, in_features, out_features), 4 * in_features, 1.0,
                                                     rngs.params()).to(dev))
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=torch.float32, device=dev))


class GroupNorm(nn.Module):
    """nnx.GroupNorm(eps=1e-6) parameter holder: scale, bias."""

    def __init__(self, num_groups, num_features, rngs=None, dtype=torch.bfloat16, param_dtype=torch.float32, device=None):
        super().__init__()
        dev = _default_device(device)
        self.num_groups, self.dtype = num_groups, dtype
        self.scale = nn.Parameter(torch.ones(num_features, dtype=torch.float32, device=dev))
        self.bias = nn.Parameter(torch.zeros(num_features, dtype=torch.float32, device=dev))


# ---------------------------------------------------------------------------------------------------------------------
# building blocks used by the Functions below (plain functions over tensors; they record what backward needs)
def _wprep(conv, which, dtype, geom, x_ld, y_ld):
    """Cached tensor-core weight image of ``conv`` (None -> generic kernel); rebuilt when the parameters change."""
    if dtype != torch.bfloat16:
        return None
    kernel = conv.kernel
    Cin, Cout = kernel.shape[3], kernel.shape[4]
    key = ("conv_wprep", which, geom, x_ld, y_ld)
    return F_.derived(kernel, key, lambda: ops.conv3d_wprep(F_.shadow(kernel, dtype), which, *geom, Cin, Cout, conv.ks,
                                                            x_ld, y_ld))


def _pad16(c):
    return (c + 15) // 16 * 16


def _padded_buffer(shape4, ld, dtype, device, kernel_pads):
    """Map with ``ld`` > C elements per voxel whose pad channels must read as zero: the tensor-core conv kernels write
    those zeros themselves (``pad_out``), the generic kernels do not."""
    shape = tuple(shape4) + (ld,)
    return torch.empty(shape, dtype=dtype, device=device) if kernel_pads else torch.zeros(shape, dtype=dtype, device=device)


def _conv_fwd(x, x_ld, Cin, conv, dtype, residual=None, out=None, out_ld=None, alloc_ld=None):
    """``alloc_ld``: produce the output at that channel pitch (> Cout, pad channels zero) in a fresh buffer."""
    w = F_.shadow(conv.kernel, dtype)
    Cout = conv.kernel.shape[4]
    if alloc_ld is not None:
        out_ld = alloc_ld
    y_ld = out_ld if (out is not None or alloc_ld is not None) else Cout
    wp = _wprep(conv, 0, dtype, tuple(x.shape[:4]), x_ld, y_ld)
    if alloc_ld is not None:
        out = _padded_buffer(x.shape[:4], alloc_ld, dtype, x.device, wp is not None and alloc_ld >= _pad16(Cout) > Cout)
    return ops.conv3d_fwd(x, w, conv.bias.detach(), conv.ks, Cin, Cout, x_ld=x_ld, residual=residual, out=out,
                          out_ld=out_ld, wprep=wp, pad_out=out is not None and out_ld >= _pad16(Cout) > Cout)


def _conv_bwd(dy, x, x_ld, Cin, conv, dtype, need_dx=True, dy_ld=None, dx_out=None, dx_ld=None, bias_done=False,
              dx_alloc_ld=None):
    """``bias_done``: the producer of ``dy`` (GroupNorm backward) already accumulated the bias gradient.
    ``dx_alloc_ld``: produce dx at that channel pitch (> Cin, pad channels zero) in a fresh buffer."""
    Cout = conv.kernel.shape[4]
    dy_ld = dy_ld or dy.shape[-1]
    gk = F_.grad_buf(conv.kernel)
    F_.wgrad_async(lambda: ops.conv3d_wgrad_accum(x, dy, gk, conv.ks, Cin, Cout, x_ld=x_ld, dy_ld=dy_ld), x, dy)
    if not bias_done:
        ops.colsum_accum(dy.reshape(-1, dy_ld)[:, :Cout], F_.grad_buf(conv.bias))
    if not need_dx:
        return None
    if dx_alloc_ld is not None:
        dx_ld = dx_alloc_ld
    o_ld = dx_ld if (dx_out is not None or dx_alloc_ld is not None) else Cin
    wp = _wprep(conv, 1, dtype, tuple(dy.shape[:4]), o_ld, dy_ld)
    if dx_alloc_ld is not None:
        dx_out = _padded_buffer(dy.shape[:4], dx_alloc_ld, dy.dtype, dy.device,
                                wp is not None and dx_alloc_ld >= _pad16(Cin) > Cin)
    return ops.conv3d_dgrad(dy, F_.shadow(conv.kernel, dtype), conv.ks, Cin, Cout, dy_ld=dy_ld, out=dx_out, out_ld=dx_ld,
                            wprep=wp, pad_out=dx_out is not None and dx_ld >= _pad16(Cin) > Cin)


class _BlockTape:
    """Saved tensors of one ConvBlock3D: conv output (pre-norm) and the GroupNorm statistics."""
    __slots__ = ("x", "x_ld", "Cin", "c", "mean", "rstd", "y", "y_ld")


def _block_fwd(x, x_ld, Cin, blk, dtype, out=None, out_ld=None):
    """ConvBlock3D: conv -> GroupNorm -> SiLU.  ``out`` lets the result land in a channel slice of a concat buffer."""
    tp = _BlockTape()
    tp.x, tp.x_ld, tp.Cin = x, x_ld, Cin
    tp.c = _conv_fwd(x, x_ld, Cin, blk.conv, dtype)
    y, tp.mean, tp.rstd = ops.groupnorm_silu_fwd(tp.c, blk.norm.scale.detach(), blk.norm.bias.detach(),
                                                 blk.norm.num_groups, out=out, out_ld=out_ld)
    tp.y, tp.y_ld = y, (out_ld if out is not None else y.shape[-1])
    return y, tp


def _block_bwd(dy, dy_ld, tp, blk, dtype, need_dx=True):
    """Returns dx with the channel stride of the block's input (``tp.x_ld``; pad channels, if any, are zero)."""
    dc = ops.groupnorm_silu_bwd(dy, dy_ld, tp.c, blk.norm.scale.detach(), blk.norm.bias.detach(), tp.mean, tp.rstd,
                                F_.grad_buf(blk.norm.scale), F_.grad_buf(blk.norm.bias), blk.norm.num_groups,
                                dx_colsum=F_.grad_buf(blk.conv.bias))
    alloc_ld = tp.x_ld if (need_dx and tp.x_ld != tp.Cin) else None
    return _conv_bwd(dc, tp.x, tp.x_ld, tp.Cin, blk.conv, dtype, need_dx, bias_done=True, dx_alloc_ld=alloc_ld)


# ---------------------------------------------------------------------------------------------------------------------
class ConvFn(Function):
    """Stand-alone nnx.Conv forward/backward (used when the holder modules are called on their own)."""

    @staticmethod
    def forward(ctx, x, kernel, bias, dtype):
        require_device()
        if x.dtype != dtype:
            x = ops.cast(x.contiguous(), dtype)
        x = x.contiguous()
        ks, Cin, Cout = tuple(kernel.shape[:3]), kernel.shape[3], kernel.shape[4]
        y = ops.conv3d_fwd(x, F_.shadow(kernel, dtype), bias.detach(), ks, Cin, Cout)
        ctx.save_for_backward(x, kernel, bias)
        ctx.dtype = dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        x, kernel, bias = ctx.saved_tensors
        ks, Cin, Cout = tuple(kernel.shape[:3]), kernel.shape[3], kernel.shape[4]
        dy = dy.contiguous()
        ops.conv3d_wgrad_accum(x, dy, F_.grad_buf(kernel), ks, Cin, Cout)
        ops.colsum_accum(dy.reshape(-1, Cout), F_.grad_buf(bias))
        dx = ops.conv3d_dgrad(dy, F_.shadow(kernel, ctx.dtype), ks, Cin, Cout) if ctx.needs_input_grad[0] else None
        F_._notify([kernel, bias])
        return dx, None, None, None


class ConvBlockFn(Function):
    @staticmethod
    def forward(ctx, x, blk, dtype, *params):
        require_device()
        if x.dtype != dtype:
            x = ops.cast(x.contiguous(), dtype)
        x = x.contiguous()
        y, tp = _block_fwd(x, x.shape[-1], x.shape[-1], blk, dtype)
        ctx.tp, ctx.blk, ctx.dtype = tp, blk, dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = _block_bwd(dy, dy.shape[-1], ctx.tp, ctx.blk, ctx.dtype)
        F_._notify(list(ctx.blk.parameters()))
        return (dx, None, None) + (None,) * 4


class ConvBlock3D(nn.Module):
    """train/unet.py:7-30."""

    def __init__(self, in_channels, out_channels, kernel_size, rngs, temporal_kernel=3, dtype=torch.bfloat16,
                 param_dtype=torch.float32, device=None):
        super().__init__()
        self.dtype = dtype
        self.conv = Conv(in_channels, out_channels, (temporal_kernel, kernel_size, kernel_size), rngs, dtype,
                         param_dtype, device=device)
        self.norm = GroupNorm(min(8, out_channels), out_channels, rngs, dtype, param_dtype, device=device)

    def forward(self, x):
        return ConvBlockFn.apply(x, self, self.dtype, self.conv.kernel, self.conv.bias, self.norm.scale, self.norm.bias)


class DownBlock3D(nn.Module):
    """train/unet.py:33-51: returns (pooled, skip)."""

    def __init__(self, in_channels, out_channels, rngs, temporal_kernel=3, dtype=torch.bfloat16,
                 param_dtype=torch.float32, device=None):
        super().__init__()
        self.conv1 = ConvBlock3D(in_channels, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype, device=device)
        self.conv2 = ConvBlock3D(out_channels, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype, device=device)

    def forward(self, x):
        x = self.conv2(self.conv1(x))
        return MaxPoolFn.apply(x), x


class UpBlock3D(nn.Module):
    """train/unet.py:54-83."""

    def __init__(self, in_channels, out_channels, rngs, temporal_kernel=3, dtype=torch.bfloat16,
                 param_dtype=torch.float32, device=None):
        super().__init__()
        self.dtype = dtype
        self.upsample = ConvTranspose(in_channels, out_channels, rngs, dtype, param_dtype, device=device)
        self.conv1 = ConvBlock3D(out_channels * 2, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype, device=device)
        self.conv2 = ConvBlock3D(out_channels, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype, device=device)

    def forward(self, x, skip):
        cat = UpsampleConcatFn.apply(x, skip, self.dtype, self.upsample.kernel, self.upsample.bias)
        return self.conv2(self.conv1(cat))


class MaxPoolFn(Function):
    @staticmethod
    def forward(ctx, x):
        require_device()
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.maxpool122_fwd(x, x.shape[-1], x.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool122_bwd(x, x.shape[-1], dy.contiguous(), None, 0, x.shape[-1])


class UpsampleConcatFn(Function):
    """cat([ConvTranspose(x), skip], channels) written straight into one buffer."""

    @staticmethod
    def forward(ctx, x, skip, dtype, kernel, bias):
        require_device()
        x, skip = x.contiguous(), skip.contiguous()
        Cout = kernel.shape[4]
        B, T, H, W, _ = x.shape
        cat = torch.empty((B, T, 2 * H, 2 * W, 2 * Cout), dtype=x.dtype, device=x.device)
        ops.convT122_fwd(x, F_.shadow(kernel, dtype), bias.detach(), Cout, cat, 2 * Cout)
        ops.copy_channels(skip, Cout, 0, cat, 2 * Cout, Cout, skip.numel() // Cout, Cout)
        ctx.save_for_backward(x, kernel, bias)
        ctx.dtype = dtype
        return cat

    @staticmethod
    def backward(ctx, dcat):
        x, kernel, bias = ctx.saved_tensors
        Cout = kernel.shape[4]
        dcat = dcat.contiguous()
        dx = ops.convT122_bwd(dcat, 2 * Cout, x, F_.shadow(kernel, ctx.dtype), F_.grad_buf(kernel), Cout)
        ops.colsum_accum(dcat.view(-1, 2 * Cout)[:, :Cout], F_.grad_buf(bias))
        dskip = torch.empty(dcat.shape[:-1] + (Cout,), dtype=dcat.dtype, device=dcat.device)
        ops.copy_channels(dcat, 2 * Cout, Cout, dskip, Cout, 0, dskip.numel() // Cout, Cout)
        F_._notify([kernel, bias])
        return dx, dskip, None, None, None


# ---------------------------------------------------------------------------------------------------------------------
class UNetFn(Function):
    """Whole U-Net (train/unet.py:155-188) + fused ``residual + unet(x)`` as a single tape."""

    @staticmethod
    def forward(ctx, x, residual, net, zero_padded, *params):
        require_device()
        dtype = net.dtype
        B, T, H, W, C0 = x.shape
        # the tensor-core conv gathers channels in blocks of 16: 12-channel maps get a 16-channel pitch (zero pads)
        C0p = _pad16(C0) if dtype == torch.bfloat16 else C0
        # ``zero_padded``: x is the [.., :C0] view of such a map already (PatchUnEmbedding produces it that way)
        pitched = (bool(zero_padded) and C0p != C0 and x.dtype == dtype and
                   tuple(x.stride()) == (T * H * W * C0p, H * W * C0p, W * C0p, C0p, 1))
        ctx.x_pitched = pitched
        if pitched:
            x = torch.as_strided(x, (B, T, H, W, C0p), x.stride(), x.storage_offset())
        else:
            if x.dtype != dtype:
                x = ops.cast(x.contiguous(), dtype)
            x = x.contiguous()
        if C0p != C0:
            if not pitched:
                xp = torch.zeros((B, T, H, W, C0p), dtype=dtype, device=x.device)
                ops.copy_channels(x, C0, 0, xp, C0p, 0, B * T * H * W, C0)
                x = xp
            cur = _conv_fwd(x, C0p, C0, net.patch_mixer, dtype, alloc_ld=C0p)
        else:
            cur = _conv_fwd(x, C0, C0, net.patch_mixer, dtype)
        tape = {"x": x, "C0": C0, "C0p": C0p}
        tape["pm"] = cur
        cur_ld, cur_c = C0p, C0
        enc_t, cats = [], []
        for enc in net.encoders:
            Cout = enc.conv1.conv.kernel.shape[4]
            a1, t1 = _block_fwd(cur, cur_ld, cur_c, enc.conv1, dtype)
            Bc, Tc, Hc, Wc = a1.shape[:4]
            cat = torch.empty((Bc, Tc, Hc, Wc, 2 * Cout), dtype=dtype, device=x.device)
            skip_view = cat[..., Cout:]                                    # skip lives in the concat buffer
            _, t2 = _block_fwd(a1, Cout, Cout, enc.conv2, dtype, out=skip_view, out_ld=2 * Cout)
            pooled = ops.maxpool122_fwd(skip_view, 2 * Cout, Cout)
            enc_t.append((t1, t2, Cout))
            cats.append(cat)
            cur, cur_ld, cur_c = pooled, Cout, Cout
        _, tb1 = _block_fwd(cur, cur_ld, cur_c, net.bottleneck1, dtype)
        Cb = net.bottleneck1.conv.kernel.shape[4]
        cur, tb2 = _block_fwd(tb1.y, Cb, Cb, net.bottleneck2, dtype)
        cur_c = Cb
        dec_t = []
        for dec, cat, (_, _, Cs) in zip(net.decoders, reversed(cats), reversed(enc_t)):
            Cout = dec.upsample.kernel.shape[4]
            assert Cout == Cs
            up_in = cur
            ops.convT122_fwd(up_in, F_.shadow(dec.upsample.kernel, dtype), dec.upsample.bias.detach(), Cout, cat, 2 * Cout)
            a1, t1 = _block_fwd(cat, 2 * Cout, 2 * Cout, dec.conv1, dtype)
            cur, t2 = _block_fwd(a1, Cout, Cout, dec.conv2, dtype)
            dec_t.append((up_in, cat, t1, t2, Cout))
            cur_c = Cout
        tape["final_in"] = cur
        out = _conv_fwd(cur, cur_c, cur_c, net.final_conv, dtype, residual=residual.contiguous() if residual is not None else None)
        tape.update(enc=enc_t, cats=cats, bott=(tb1, tb2), dec=dec_t)
        ctx.tape, ctx.net = tape, net
        ctx.has_res = residual is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        tape, net = ctx.tape, ctx.net
        dtype = net.dtype
        dout = dout.contiguous()
        fin = tape["final_in"]
        dcur = _conv_bwd(dout, fin, fin.shape[-1], fin.shape[-1], net.final_conv, dtype)
        dskips = []
        for dec, (up_in, cat, t1, t2, Cout) in zip(reversed(list(net.decoders)), reversed(tape["dec"])):
            da1 = _block_bwd(dcur, Cout, t2, dec.conv2, dtype)
            dcat = _block_bwd(da1, Cout, t1, dec.conv1, dtype)                      # [.., 2*Cout]
            dprev = ops.convT122_bwd(dcat, 2 * Cout, up_in, F_.shadow(dec.upsample.kernel, dtype),
                                     F_.grad_buf(dec.upsample.kernel), Cout)
            ops.colsum_accum(dcat.view(-1, 2 * Cout)[:, :Cout], F_.grad_buf(dec.upsample.bias))
            dskips.append(dcat)                                                     # skip grad = dcat[..., Cout:]
            dcur = dprev
        tb1, tb2 = tape["bott"]
        Cb = net.bottleneck1.conv.kernel.shape[4]
        d1 = _block_bwd(dcur, Cb, tb2, net.bottleneck2, dtype)
        dcur = _block_bwd(d1, Cb, tb1, net.bottleneck1, dtype)
        # dskips was filled shallowest level first (decoder backward runs last decoder first); the encoders unwind
        # deepest level first, hence reversed(dskips)
        for enc, (t1, t2, Cout), cat, dcat in zip(reversed(list(net.encoders)), reversed(tape["enc"]),
                                                  reversed(tape["cats"]), reversed(dskips)):
            skip_view = cat[..., Cout:]
            da2 = ops.maxpool122_bwd(skip_view, 2 * Cout, dcur, dcat[..., Cout:], 2 * Cout, Cout)
            da1 = _block_bwd(da2, Cout, t2, enc.conv2, dtype)
            dcur = _block_bwd(da1, Cout, t1, enc.conv1, dtype)
        x, C0, C0p = tape["x"], tape["C0"], tape["C0p"]
        need_dx = ctx.needs_input_grad[0]
        if C0p != C0:
            dxp = _conv_bwd(dcur, x, C0p, C0, net.patch_mixer, dtype, need_dx=need_dx, dy_ld=C0p,
                            dx_alloc_ld=C0p if need_dx else None)
            dx = None
            if need_dx and ctx.x_pitched:
                dx = dxp[..., :C0]                     # the consumer (UnembedFn.backward) reads it at this pitch
            elif need_dx:
                dx = torch.empty(tuple(x.shape[:4]) + (C0,), dtype=x.dtype, device=x.device)
                ops.copy_channels(dxp, C0p, 0, dx, C0, 0, dx.numel() // C0, C0)
        else:
            dx = _conv_bwd(dcur, x, C0, C0, net.patch_mixer, dtype, need_dx=need_dx)
        F_._notify(list(net.parameters()))
        ctx.tape = None
        return (dx, dout if ctx.has_res else None, None, None) + (None,) * (len(list(net.parameters())))


class UNet(nn.Module):
    """train/unet.py:86-188."""

    def __init__(self, channels, base_features=32, num_levels=3, out_features=3, rngs=None, temporal_kernel=3,
                 dtype=torch.bfloat16, param_dtype=torch.float32, device=None):
        super().__init__()
        self.num_levels, self.dtype = num_levels, dtype
        self.patch_mixer = Conv(channels, channels, (temporal_kernel, 7, 7), rngs, dtype, param_dtype, device=device)
        self.encoders = nn.ModuleList()
        in_ch = channels
        for i in range(num_levels):
            out_ch = base_features * (2 ** i)
            self.encoders.append(DownBlock3D(in_ch, out_ch, rngs, temporal_kernel, dtype, param_dtype, device=device))
            in_ch = out_ch
        bott = base_features * (2 ** num_levels)
        self.bottleneck1 = ConvBlock3D(in_ch, bott, 3, rngs, temporal_kernel, dtype, param_dtype, device=device)
        self.bottleneck2 = ConvBlock3D(bott, bott, 3, rngs, temporal_kernel, dtype, param_dtype, device=device)
        self.decoders = nn.ModuleList()
        in_ch = bott
        for i in range(num_levels - 1, -1, -1):
            out_ch = base_features * (2 ** i)
            self.decoders.append(UpBlock3D(in_ch, out_ch, rngs, temporal_kernel, dtype, param_dtype, device=device))
            in_ch = out_ch
        self.final_conv = Conv(base_features, out_features, (1, 1, 1), rngs, dtype, param_dtype, zero_init=True,
                               device=device)

    def forward(self, x, residual=None, zero_padded=False):
        """``residual`` (optional, [b,t,H,W,out]) is added to the output inside final_conv's epilogue.
        ``zero_padded``: x is the [.., :C] view of a map with ceil16(C) elements per voxel whose pad channels are zero
        (what PatchUnEmbedding returns in bf16): it is consumed in place."""
        return UNetFn.apply(x, residual, self, zero_padded, *self.parameters())
