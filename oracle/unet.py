"""Oracle restatement of train/unet.py:7-188 (CPU torch, TEST INFRASTRUCTURE ONLY)."""
import torch
from torch import nn

from . import nn as onn


class ConvBlock3D(nn.Module):
    """train/unet.py:7-30 -- Conv(kt,k,k 'SAME') -> GroupNorm(min(8,C)) -> SiLU."""

    def __init__(self, in_channels, out_channels, kernel_size, rngs, temporal_kernel=3,
                 dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.conv = onn.Conv(in_channels, out_channels, (temporal_kernel, kernel_size, kernel_size), rngs, dtype, param_dtype)
        self.norm = onn.GroupNorm(min(8, out_channels), out_channels, rngs, dtype, param_dtype)

    def forward(self, x):
        return onn.silu(self.norm(self.conv(x)))


class DownBlock3D(nn.Module):
    """train/unet.py:33-51."""

    def __init__(self, in_channels, out_channels, rngs, temporal_kernel=3, dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.conv1 = ConvBlock3D(in_channels, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype)
        self.conv2 = ConvBlock3D(out_channels, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype)

    def forward(self, x):
        x = self.conv2(self.conv1(x))
        return onn.max_pool_122(x), x


class UpBlock3D(nn.Module):
    """train/unet.py:54-83."""

    def __init__(self, in_channels, out_channels, rngs, temporal_kernel=3, dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.upsample = onn.ConvTranspose122(in_channels, out_channels, rngs, dtype, param_dtype)
        self.conv1 = ConvBlock3D(out_channels * 2, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype)
        self.conv2 = ConvBlock3D(out_channels, out_channels, 3, rngs, temporal_kernel, dtype, param_dtype)

    def forward(self, x, skip):
        x = torch.cat([self.upsample(x), skip], dim=-1)
        return self.conv2(self.conv1(x))


class UNet(nn.Module):
    """train/unet.py:86-188 (final_conv kernel zero-initialised, :144-153)."""

    def __init__(self, channels, base_features=32, num_levels=3, out_features=3, rngs=None, temporal_kernel=3,
                 dtype=torch.float32, param_dtype=torch.float32):
        super().__init__()
        self.num_levels, self.dtype = num_levels, dtype
        self.patch_mixer = onn.Conv(channels, channels, (temporal_kernel, 7, 7), rngs, dtype, param_dtype)
        self.encoders = nn.ModuleList()
        in_ch = channels
        for i in range(num_levels):
            out_ch = base_features * (2 ** i)
            self.encoders.append(DownBlock3D(in_ch, out_ch, rngs, temporal_kernel, dtype, param_dtype))
            in_ch = out_ch
        bott = base_features * (2 ** num_levels)
        self.bottleneck1 = ConvBlock3D(in_ch, bott, 3, rngs, temporal_kernel, dtype, param_dtype)
        self.bottleneck2 = ConvBlock3D(bott, bott, 3, rngs, temporal_kernel, dtype, param_dtype)
        self.decoders = nn.ModuleList()
        in_ch = bott
        for i in range(num_levels - 1, -1, -1):
            out_ch = base_features * (2 ** i)
            self.decoders.append(UpBlock3D(in_ch, out_ch, rngs, temporal_kernel, dtype, param_dtype))
            in_ch = out_ch
        self.final_conv = onn.Conv(base_features, out_features, (1, 1, 1), rngs, dtype, param_dtype, zero_init=True)

    def forward(self, x):
        x = self.patch_mixer(x.to(self.dtype))
        skips = []
        for enc in self.encoders:
            x, skip = enc(x)
            skips.append(skip)
        x = self.bottleneck2(self.bottleneck1(x))
        for dec, skip in zip(self.decoders, reversed(skips)):
            x = dec(x, skip)
        return self.final_conv(x)
