"""B200-native RL training loss of the reference's production loop.

``loss_fn`` mirrors train/rl_nonadversarial.py:100-186 over the 6-tuple of ``rl_model.VideoVAE`` (signature, loss
terms, returned aux keys); ``train_step`` / ``eval_step`` mirror :188-209 (mask plumbing + value_and_grad; eval_step
keeps the reference's ``train=True``), leaving the optimizer update to the caller.

The video-sized work (per-sample masked MSE / MAE / KL and their gradients) runs in the CUDA kernels behind
``vvae_recon_loss_per_sample_fwd`` / ``vvae_kl_per_sample_fwd`` / ``vvae_recon_loss_bwd`` / ``vvae_kl_bwd``; the pairwise
advantage and trajectory-probability terms are a few ``(b, 2, t)`` tensors.  The VGG perceptual term (:125) is the
caller's ``perceptual_loss_fn(vgg_params, reconstruction, video) -> [b]`` (pretrained weights are not available
offline); ``None`` drops it.
"""
import torch
from torch.autograd import Function

from . import ops
from ._ffi import require_device
from .rl_model import repeat2

DEFAULT_HPARAMS = {  # rl_nonadversarial.py:47-57, 255-263
    "gamma1": 0.2,
    "gamma2": 0.001,
    "gamma3": 0.1,
    "gamma4": 0.05,
    "max_compression_rate": 2,
    "magnify_negatives_rate": 100,
    "rl_loss_weight": 0.01,
}


def per_sample_mean(x):
    """rl_nonadversarial.py:59-60."""
    return x.mean(dim=tuple(range(1, x.ndim)))


def magnify_negatives(x, rate):
    """rl_nonadversarial.py:70-72."""
    return torch.where(x < 0, x * rate, x)


class PerSampleLossFn(Function):
    """Per-sample masked reconstruction and KL terms (rl_nonadversarial.py:114-121, 145-147).

    Returns (MSE_b + gamma4 * MAE_b, KL_b, MSE_b, MAE_b), each fp32 [B]; the first two are differentiable w.r.t.
    (recon) and (logvar, mean) with an arbitrary per-sample upstream gradient, which the backward folds into the
    per-sample 1/len weights the kernels already take."""

    @staticmethod
    def forward(ctx, video, recon, logvar, mean, mask_bt, gamma4):
        require_device()
        B, T = mask_bt.shape
        dev = recon.device
        m = mask_bt.to(torch.float32).contiguous()
        inv_len = (1.0 / torch.clamp(m.sum(dim=1), min=1.0)).contiguous()
        frame_w = (m * inv_len[:, None]).contiguous()
        video, recon = video.contiguous(), recon.contiguous()
        mean, logvar = mean.contiguous(), logvar.contiguous()
        sums = ops.zeros_f32((B, 2), dev)
        ops.recon_loss_per_sample_fwd(video, recon, m.reshape(-1), inv_len, sums)
        kl_sum = ops.zeros_f32((B,), dev)
        ops.kl_per_sample_fwd(mean, logvar, frame_w.reshape(-1), kl_sum, mean.shape[2])
        per_frame = video.numel() // (B * T)
        mse, mae = sums[:, 0] / float(per_frame), sums[:, 1] / float(per_frame)
        kl = kl_sum / float(mean.numel() // B)
        ctx.save_for_backward(video, recon, logvar, mean, m, inv_len, frame_w)
        ctx.gamma4, ctx.per_frame = float(gamma4), per_frame
        ctx.mark_non_differentiable(mse, mae)
        return mse + float(gamma4) * mae, kl, mse, mae

    @staticmethod
    def backward(ctx, g_rec, g_kl, *unused):
        video, recon, logvar, mean, m, inv_len, frame_w = ctx.saved_tensors
        B = m.shape[0]
        drecon = dmean = dlogvar = None
        if g_rec is not None and ctx.needs_input_grad[1]:
            w = (inv_len * g_rec.to(torch.float32)).contiguous()
            drecon = ops.recon_loss_bwd(video, recon, m.reshape(-1), w, 1.0, ctx.gamma4, 1.0 / ctx.per_frame)
        if g_kl is not None and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3]):
            fw = (frame_w * g_kl.to(torch.float32)[:, None]).reshape(-1).contiguous()
            dmean, dlogvar = ops.kl_bwd(mean, logvar, fw, 1.0 / float(mean.numel() // B), mean.shape[2])
        return None, drecon, dlogvar, dmean, None, None


def loss_terms(video, reconstruction, selection, selection_mask, logvar, mean, original_mask, hparams,
               perceptual_loss_fn=None, vgg_params=None):
    """rl_nonadversarial.py:104-186, everything after the model call."""
    output_mask = repeat2(original_mask)                                           # :104
    m = output_mask.to(torch.float32)
    seq = torch.clamp(m.sum(dim=1, keepdim=True), min=1.0)                         # :105-106
    video2 = repeat2(video)                                                        # :110
    rec_ps, kl_loss, per_sample_error, per_sample_mae = PerSampleLossFn.apply(
        video2, reconstruction, logvar, mean, output_mask, hparams["gamma4"])
    if perceptual_loss_fn is not None:
        perceptual = perceptual_loss_fn(vgg_params, reconstruction, video2.to(reconstruction.dtype)).to(torch.float32)
    else:
        perceptual = torch.zeros_like(rec_ps)
    t = m.shape[1]
    density = (selection_mask.reshape(-1, t).to(torch.float32) * m).sum(dim=1, keepdim=True) / seq     # :130-133
    diff = density - (1.0 / hparams["max_compression_rate"])
    selection_loss = per_sample_mean(torch.square(magnify_negatives(diff, hparams["magnify_negatives_rate"])))
    per_sample_loss = (rec_ps + hparams["gamma3"] * perceptual + hparams["gamma1"] * selection_loss
                       + hparams["gamma2"] * kl_loss)                              # :149
    pairs = per_sample_loss.reshape(-1, 2)
    means = pairs.mean(dim=1, keepdim=True)
    stds = pairs.std(dim=1, unbiased=False, keepdim=True) + 1e-6
    disadvantages = ((pairs - means) / stds).detach()                              # :153, :173
    actions = selection_mask.reshape(-1, 2, t).to(torch.float32)
    sel = selection.reshape(-1, 2, t).to(torch.float32)
    raw_probs = torch.clamp(torch.abs(sel + actions - 1), 1e-6, 1.0 - 1e-6)       # :163
    rl_mask = output_mask.reshape(-1, 2, t).to(torch.bool)
    one = torch.ones_like(raw_probs)
    # :164-171  prod_t(p_t) with every factor p_t = raw/stop_grad(raw) == 1: value 1, d/dp_t = 1.  Written as
    # 1 + sum_t(p_t - 1), which has exactly that value and gradient; torch's prod backward looks for zeros with a host
    # read, which a CUDA-graph capture forbids.
    ratio = torch.where(rl_mask, raw_probs / raw_probs.detach(), one)
    probs = 1.0 + (ratio - 1.0).sum(dim=2, keepdim=True)
    raw_traj = torch.where(rl_mask, raw_probs, one).detach().prod(dim=2, keepdim=True)   # :168-169 (logged only)
    rl_loss = probs * disadvantages[:, :, None]
    loss = per_sample_loss.mean() + rl_loss.mean() * hparams["rl_loss_weight"]     # :174
    return loss, {
        "MSE": per_sample_error.mean(), "perceptual_loss": perceptual.mean(), "selection_loss": selection_loss.mean(),
        "kl_loss": kl_loss.mean(), "kept_frame_density": density.mean(), "mean_trajectory_prob": raw_traj.mean(),
        "rl_loss": rl_loss.mean(), "per_sample_MAE": per_sample_mae.mean(), "per_sample_loss": per_sample_loss,
    }


def loss_fn(model, video, mask, original_mask, rngs, hparams=None, perceptual_loss_fn=None, vgg_params=None, train=True,
            noise=None, bernoulli_u=None):
    hparams = DEFAULT_HPARAMS if hparams is None else hparams
    reconstruction, compressed, selection, selection_mask, logvar, mean = model(
        video, mask, rngs, train=train, noise=noise, bernoulli_u=bernoulli_u)
    loss, aux = loss_terms(video, reconstruction, selection, selection_mask, logvar, mean, original_mask, hparams,
                           perceptual_loss_fn, vgg_params)
    aux.update(reconstruction=reconstruction, selection=selection, selection_mask=selection_mask)
    return loss, aux


def train_step(model, video, mask_bt, hparams, rngs, perceptual_loss_fn=None, vgg_params=None, noise=None,
               bernoulli_u=None):
    """rl_nonadversarial.py:188-198 without ``optimizer.update``: gradients are left accumulated (fp32)."""
    loss, aux = loss_fn(model, video, mask_bt[:, None, None, :], mask_bt, rngs, hparams, perceptual_loss_fn, vgg_params,
                        train=True, noise=noise, bernoulli_u=bernoulli_u)
    loss.backward()
    return loss, aux


def eval_step(model, video, mask_bt, hparams, rngs, perceptual_loss_fn=None, vgg_params=None, noise=None,
              bernoulli_u=None):
    """rl_nonadversarial.py:200-209 (the reference evaluates with train=True)."""
    with torch.no_grad():
        return loss_fn(model, video, mask_bt[:, None, None, :], mask_bt, rngs, hparams, perceptual_loss_fn, vgg_params,
                       train=True, noise=noise, bernoulli_u=bernoulli_u)
