"""Oracle restatement of the VGG-16 perceptual loss (CPU torch, TEST INFRASTRUCTURE ONLY).  Parity unpinned.

Follows train/vgg_tests.py:8-68 (load_vgg, PERCEPTUAL_LAYERS, get_adversarial_perceptual_loss_fn,
get_perceptual_loss_fn).  The network itself is the third-party ``flaxmodels==0.1.3`` ``VGG16(output='activations',
include_head=False, normalize=True)`` (claude_distributed/requirements.txt), absent from /root/reference; its published
definition up to the deepest layer the loss reads is restated here: ImageNet normalisation
``(x - [0.485, 0.456, 0.406]) / [0.229, 0.224, 0.225]``, then conv1_1 (3->64), ReLU, conv1_2 (64->64), ReLU,
2x2/stride-2 max pool, conv2_1 (64->128), ReLU; all convolutions 3x3, stride 1, 'SAME', with bias; kernels in Flax HWIO
layout.  Pretrained weights need network access: parameters are random unless loaded from a flaxmodels parameter tree.
"""
import torch
import torch.nn.functional as F
from torch import nn

PERCEPTUAL_LAYERS = ("relu1_1", "relu1_2", "relu2_1")      # vgg_tests.py:36
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_LAYERS = (("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128))


class VGG16Features(nn.Module):
    def __init__(self, rngs, normalize=True, dtype=torch.float32):
        super().__init__()
        key = rngs.params()
        self.normalize, self.dtype = normalize, dtype
        for name, cin, cout in _LAYERS:
            k = torch.randn(3, 3, cin, cout, generator=key, dtype=torch.float32) * (2.0 / (9 * cin)) ** 0.5
            b = torch.randn(cout, generator=key, dtype=torch.float32) * 0.05
            setattr(self, name + "_kernel", nn.Parameter(k, requires_grad=False))
            setattr(self, name + "_bias", nn.Parameter(b, requires_grad=False))

    def _conv(self, x, name):
        w = getattr(self, name + "_kernel").to(x.dtype).permute(3, 2, 0, 1)          # HWIO -> OIHW
        y = F.conv2d(x.permute(0, 3, 1, 2), w, getattr(self, name + "_bias").to(x.dtype), padding=1)
        return y.permute(0, 2, 3, 1)

    def forward(self, x):
        """x [n, H, W, 3] in [0, 1] -> {'relu1_1', 'relu1_2', 'relu2_1'} (channels last)."""
        x = x.to(self.dtype)
        if self.normalize:
            x = (x - torch.tensor(IMAGENET_MEAN, dtype=x.dtype)) / torch.tensor(IMAGENET_STD, dtype=x.dtype)
        acts = {}
        acts["relu1_1"] = torch.relu(self._conv(x, "conv1_1"))
        acts["relu1_2"] = torch.relu(self._conv(acts["relu1_1"], "conv1_2"))
        pooled = F.max_pool2d(acts["relu1_2"].permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        acts["relu2_1"] = torch.relu(self._conv(pooled, "conv2_1"))
        return acts


def get_adversarial_perceptual_loss_fn(model):
    """vgg_tests.py:38-68: (params, x, target) -> [b]; per-frame sum of the three layers' mean squared feature
    differences, averaged over time.  ``params`` is accepted for signature parity (the weights live in ``model``)."""
    def perceptual_loss(params, x, target):
        b, t = x.shape[:2]
        fx = model(x.reshape((b * t,) + tuple(x.shape[2:])))
        ft = model(target.reshape((b * t,) + tuple(target.shape[2:])))
        per_frame = sum(((fx[k] - ft[k]) ** 2).float().mean(dim=(1, 2, 3)) for k in PERCEPTUAL_LAYERS)
        return per_frame.reshape(b, t).mean(dim=-1)
    return perceptual_loss


def get_perceptual_loss_fn(model):
    """vgg_tests.py:70-97: the scalar form (global means)."""
    def perceptual_loss(params, x, target):
        b, t = x.shape[:2]
        fx = model(x.reshape((b * t,) + tuple(x.shape[2:])))
        ft = model(target.reshape((b * t,) + tuple(target.shape[2:])))
        return sum(((fx[k] - ft[k]) ** 2).float().mean() for k in PERCEPTUAL_LAYERS)
    return perceptual_loss
