"""B200-native VGG-16 perceptual loss: drop-in for train/vgg_tests.py.

``load_vgg`` / ``get_adversarial_perceptual_loss_fn`` / ``get_perceptual_loss_fn`` mirror vgg_tests.py:8-97.  The
feature extractor is flaxmodels 0.1.3 ``VGG16(output='activations', include_head=False, normalize=True)`` restricted
to what the loss reads (relu1_1, relu1_2, relu2_1).  Its 3x3 convolutions run on the same conv3d kernels as the U-Net
(kt = 1: every frame is an image; the tcgen05 implicit-GEMM path in bf16), around them ``vvae_vgg_preprocess_*``
(ImageNet normalisation into a zero-padded 16-channel tensor), ``vvae_relu_*``, ``vvae_maxpool122_*`` and the
per-sample squared-error kernels of the reconstruction loss.  Only the gradient w.r.t. the reconstruction is produced
(the reference differentiates the VAE, not the VGG weights).

ImageNet weights are not available offline: parameters are random unless ``load_flax_params`` is given a flaxmodels
parameter tree (``{'params': {'conv1_1': {'kernel', 'bias'}, ...}}``, HWIO kernels -- stored here unchanged).
"""
import torch
from torch import nn
from torch.autograd import Function

from . import functional as F_
from . import ops
from ._ffi import require_device
from .layers import _default_device

PERCEPTUAL_LAYERS = ("relu1_1", "relu1_2", "relu2_1")      # vgg_tests.py:36
_LAYERS = (("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128))
_KS = (1, 3, 3)
_PAD_IN = 16                                                # RGB padded to the tensor-core conv's channel granule


class VGG16Features(nn.Module):
    def __init__(self, rngs, dtype=torch.bfloat16, device=None):
        super().__init__()
        key = rngs.params()
        self.dtype = dtype
        dev = _default_device(device)
        for name, cin, cout in _LAYERS:
            k = torch.randn(3, 3, cin, cout, generator=key, dtype=torch.float32) * (2.0 / (9 * cin)) ** 0.5
            b = torch.randn(cout, generator=key, dtype=torch.float32) * 0.05
            setattr(self, name + "_kernel", nn.Parameter(k.to(dev), requires_grad=False))
            setattr(self, name + "_bias", nn.Parameter(b.to(dev), requires_grad=False))
        self._wprep = {}

    def load_flax_params(self, tree):
        params = tree.get("params", tree)
        with torch.no_grad():
            for name, _, _ in _LAYERS:
                node = params[name]
                getattr(self, name + "_kernel").copy_(torch.as_tensor(node["kernel"], dtype=torch.float32))
                getattr(self, name + "_bias").copy_(torch.as_tensor(node["bias"], dtype=torch.float32))
        self._wprep.clear()
        F_.invalidate_shadows()

    # -- conv helpers: weights as [1,3,3,Cin,Cout] in the compute dtype, tensor-core weight images cached per geometry
    def _w(self, name):
        k = getattr(self, name + "_kernel")
        return F_.shadow(k, self.dtype).reshape((1,) + tuple(k.shape))

    def _img(self, name, which, geom, cin, cout, x_ld, y_ld):
        key = (name, which, geom, x_ld, y_ld)
        if key not in self._wprep:
            self._wprep[key] = ops.conv3d_wprep(self._w(name), which, *geom, cin, cout, _KS, x_ld, y_ld)
        return self._wprep[key]

    def conv_relu(self, x, x_ld, name, cin, cout):
        geom = tuple(x.shape[:4])
        y = ops.conv3d_fwd(x, self._w(name), getattr(self, name + "_bias").detach(), _KS, cin, cout, x_ld=x_ld,
                           wprep=self._img(name, 0, geom, cin, cout, x_ld, cout))
        return ops.relu_(y)

    def conv_dgrad(self, dy, name, cin, cout, out_ld=None):
        geom = tuple(dy.shape[:4])
        o_ld = out_ld or cin
        out = None
        if out_ld is not None:
            out = torch.empty(geom + (out_ld,), dtype=dy.dtype, device=dy.device)
        return ops.conv3d_dgrad(dy, self._w(name), _KS, cin, cout, out=out, out_ld=out_ld,
                                wprep=self._img(name, 1, geom, cin, cout, o_ld, cout),
                                pad_out=out_ld is not None and out_ld > cin)

    def features(self, x):
        """x [b,t,H,W,3] in [0,1] -> (relu1_1, relu1_2, relu2_1), channels last, compute dtype."""
        require_device()
        p0 = ops.vgg_preprocess_fwd(x, self.dtype, _PAD_IN)
        r11 = self.conv_relu(p0, _PAD_IN, "conv1_1", 3, 64)
        r12 = self.conv_relu(r11, 64, "conv1_2", 64, 64)
        pooled = ops.maxpool122_fwd(r12, 64, 64)
        r21 = self.conv_relu(pooled, 64, "conv2_1", 64, 128)
        return r11, r12, r21

    def forward(self, x):
        """flaxmodels-style activations dict for frames x [n,H,W,3] (no gradient)."""
        with torch.no_grad():
            r = self.features(x[:, None])
        return {k: v[:, 0] for k, v in zip(PERCEPTUAL_LAYERS, r)}


def _frame_sq_err(fx, ft):
    """Per-frame mean squared difference of two feature tensors [b,t,h,w,c] -> fp32 [b*t]."""
    b, t = fx.shape[:2]
    n = b * t
    per = fx.numel() // n
    ones = torch.ones(n, dtype=torch.float32, device=fx.device)
    out = ops.zeros_f32((n, 2), fx.device)
    ops.recon_loss_per_sample_fwd(ft.reshape(n, 1, per), fx.reshape(n, 1, per), ones, ones, out)
    return out[:, 0] / float(per)


def _frame_sq_err_bwd(fx, ft, w_frame):
    """d/dfx of sum_n w_frame[n] * mean((fx_n - ft_n)^2)."""
    b, t = fx.shape[:2]
    n = b * t
    per = fx.numel() // n
    ones = torch.ones(n, dtype=torch.float32, device=fx.device)
    d = ops.recon_loss_bwd(ft.reshape(n, 1, per), fx.reshape(n, 1, per), ones, w_frame.contiguous(), 1.0, 0.0, 1.0 / per)
    return d.reshape(fx.shape)


class VggPerceptualFn(Function):
    """Per-frame perceptual distance sum_l mean((phi_l(x) - phi_l(target))^2), fp32 [b*t]; gradient w.r.t. x only."""

    @staticmethod
    def forward(ctx, x, target, vgg):
        with torch.no_grad():
            ft = vgg.features(target)
            fx = vgg.features(x)
            per_frame = sum(_frame_sq_err(a, b_) for a, b_ in zip(fx, ft))
        ctx.vgg, ctx.fx, ctx.ft, ctx.x_dtype = vgg, fx, ft, x.dtype
        return per_frame

    @staticmethod
    def backward(ctx, g):
        vgg, (r11, r12, r21), (t11, t12, t21) = ctx.vgg, ctx.fx, ctx.ft
        w = g.detach().to(torch.float32).reshape(-1)
        d21 = ops.relu_bwd(_frame_sq_err_bwd(r21, t21, w), r21)
        dpool = vgg.conv_dgrad(d21, "conv2_1", 64, 128)
        # max-pool backward adds the relu1_2 term of the loss as its "skip" gradient
        d12 = ops.maxpool122_bwd(r12, 64, dpool, _frame_sq_err_bwd(r12, t12, w), 64, 64)
        d12 = ops.relu_bwd(d12, r12)
        d11 = vgg.conv_dgrad(d12, "conv1_2", 64, 64)
        d11.add_(_frame_sq_err_bwd(r11, t11, w))                       # two consumers of relu1_1: sum of gradients
        d11 = ops.relu_bwd(d11, r11)
        dp0 = vgg.conv_dgrad(d11, "conv1_1", 3, 64, out_ld=_PAD_IN)
        dx = ops.vgg_preprocess_bwd(dp0)
        ctx.fx = ctx.ft = None
        return dx.to(ctx.x_dtype), None, None


def load_vgg(rngs=None, dtype=torch.bfloat16, device=None, flax_params=None):
    """vgg_tests.py:8-33.  Returns (model, params); ``params`` is kept for signature parity (weights live in model)."""
    from .rng import Rngs
    model = VGG16Features(rngs if rngs is not None else Rngs(0), dtype=dtype, device=device)
    if flax_params is not None:
        model.load_flax_params(flax_params)
    return model, None


def get_adversarial_perceptual_loss_fn(model):
    """vgg_tests.py:38-68: (params, x, target) -> [b] (per-frame distances averaged over time)."""
    def perceptual_loss(params, x, target):
        b, t = x.shape[:2]
        return VggPerceptualFn.apply(x, target, model).reshape(b, t).mean(dim=-1)
    return perceptual_loss


def get_perceptual_loss_fn(model):
    """vgg_tests.py:70-97: scalar form.  Frames have equal sizes, so the global mean is the mean of per-frame means."""
    def perceptual_loss(params, x, target):
        return VggPerceptualFn.apply(x, target, model).mean()
    return perceptual_loss
