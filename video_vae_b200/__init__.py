"""video_vae_b200 -- B200-native (sm_100a) implementation of the floatingtrees/video-VAE encoder/decoder hot path.

Python is the host language (as in the reference); every kernel is hand-written CUDA reached through the C ABI in
include/vvae.h (video_vae_b200/libvvae.so).  Importing this package REQUIRES the built library and there is no CPU
or PyTorch-op fallback: calls fail loudly without a compute-capability-10.x device.
"""
from . import _ffi  # noqa: F401  (raises ImportError if libvvae.so is missing)
from .layers import (MLP, Attention, FactoredAttention, GumbelSigmoidSTE, PatchEmbedding, PatchUnEmbedding,  # noqa: F401
                     RotaryEmbedding, rotate_half, round_ste)
from .losses import DEFAULT_HPARAMS, expand_mask, loss_fn, train_step  # noqa: F401
from .model import Decoder, Encoder, VideoVAE  # noqa: F401
from .rng import Rngs  # noqa: F401
from .unet import UNet, ConvBlock3D, DownBlock3D, UpBlock3D  # noqa: F401

__all__ = ["VideoVAE", "Encoder", "Decoder", "UNet", "ConvBlock3D", "DownBlock3D", "UpBlock3D", "FactoredAttention",
           "Attention", "MLP", "PatchEmbedding", "PatchUnEmbedding", "RotaryEmbedding", "GumbelSigmoidSTE", "round_ste",
           "rotate_half", "Rngs", "loss_fn", "train_step", "expand_mask", "DEFAULT_HPARAMS"]
