"""Oracle restatement of the RL training loss (CPU torch, TEST INFRASTRUCTURE ONLY).  Pinned against the reference's own
loss_fn executed on oracle/jaxshim (tests/golden/refshim_rlvae_*.npz, tests/test_jax_golden.py); not against real JAX.

Follows train/rl_nonadversarial.py:59-60 (per_sample_mean), :70-72 (magnify_negatives), :100-186 (loss_fn) and
:188-209 (train_step / eval_step mask plumbing) over the 6-tuple of train/rl_model.py.  The VGG perceptual term
(:125, weights need network access) is a caller-supplied ``perceptual_loss_fn(vgg_params, reconstruction, video) -> [b]``;
``None`` drops the term (gamma3 * 0).
"""
import torch

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
    """:59-60."""
    return x.mean(dim=tuple(range(1, x.ndim)))


def magnify_negatives(x, rate):
    """:70-72."""
    return torch.where(x < 0, x * rate, x)


def loss_terms(video, reconstruction, selection, selection_mask, logvar, mean, original_mask, hparams,
               perceptual_loss_fn=None, vgg_params=None):
    """:104-186, everything after the model call."""
    f32 = torch.promote_types(torch.float32, reconstruction.dtype)
    cd = reconstruction.dtype
    output_mask = original_mask.repeat_interleave(2, dim=0)                       # :104  (b 2) time
    seq = torch.clamp(output_mask.to(f32).sum(dim=1, keepdim=True), min=1.0)      # :105-106
    vm = output_mask.to(cd)[:, :, None, None, None]
    video2 = video.repeat_interleave(2, dim=0).to(cd)                             # :110
    seq5 = seq[:, :, None, None, None]

    err = (video2 - reconstruction) * vm
    ab = torch.abs(err)                                                           # :114
    frame_mae = ab.to(f32).sum(dim=1, keepdim=True).to(cd).to(f32) / seq5         # :116
    per_sample_mae = per_sample_mean(frame_mae)
    sq = torch.square(err)                                                        # :119
    frame_err = sq.to(f32).sum(dim=1, keepdim=True).to(cd).to(f32) / seq5         # :120
    per_sample_error = per_sample_mean(frame_err)

    if perceptual_loss_fn is not None:
        perceptual = perceptual_loss_fn(vgg_params, reconstruction, video2).to(f32)   # :125
    else:
        perceptual = torch.zeros_like(per_sample_error)

    km = output_mask.to(f32)[:, :, None, None]                                    # :127
    sel_sum = (selection_mask.to(f32) * km).sum(dim=(1, 2, 3))[:, None]           # :130
    density = sel_sum / seq                                                       # :133
    diff = density - (1.0 / hparams["max_compression_rate"])                      # :139
    selection_loss = per_sample_mean(torch.square(magnify_negatives(diff, hparams["magnify_negatives_rate"])))

    kl_el = 0.5 * (torch.exp(logvar) - 1 - logvar + torch.square(mean))           # :146
    kl = (kl_el * output_mask.to(kl_el.dtype)[:, :, None, None]).to(f32) / seq[:, :, None, None]
    kl_loss = per_sample_mean(kl)

    per_sample_loss = (per_sample_error + hparams["gamma3"] * perceptual + hparams["gamma1"] * selection_loss
                       + hparams["gamma2"] * kl_loss + hparams["gamma4"] * per_sample_mae)    # :149
    pairs = per_sample_loss.reshape(-1, 2)                                        # :150
    means = pairs.mean(dim=1, keepdim=True)
    stds = pairs.std(dim=1, unbiased=False, keepdim=True) + 1e-6                  # jnp.std: ddof 0
    disadvantages = (pairs - means) / stds                                        # :153
    t = output_mask.shape[1]
    actions = selection_mask.to(f32).reshape(-1, 2, t)                            # :154
    sel = selection.to(f32).reshape(-1, 2, t)                                     # :157
    raw_probs = torch.clamp(torch.abs(sel + actions - 1), 1e-6, 1.0 - 1e-6)      # :163
    probs = raw_probs / raw_probs.detach()                                        # :164
    rl_mask = output_mask.reshape(-1, 2, t).to(torch.bool)
    probs = torch.where(rl_mask, probs, torch.ones_like(probs))                   # :166
    raw_masked = torch.where(rl_mask, raw_probs, torch.ones_like(raw_probs))
    raw_traj = raw_masked.prod(dim=2, keepdim=True)                               # :169
    probs = probs.prod(dim=2, keepdim=True)                                       # :171
    rl_loss = probs * disadvantages.detach()[:, :, None]                          # :173
    loss = per_sample_loss.mean() + rl_loss.mean() * hparams["rl_loss_weight"]    # :174
    return loss, {
        "MSE": per_sample_error.mean(), "perceptual_loss": perceptual.mean(), "selection_loss": selection_loss.mean(),
        "kl_loss": kl_loss.mean(), "kept_frame_density": density.mean(), "mean_trajectory_prob": raw_traj.mean(),
        "rl_loss": rl_loss.mean(), "per_sample_MAE": per_sample_mae.mean(), "per_sample_loss": per_sample_loss,
    }


def loss_fn(model, video, mask, original_mask, rngs, hparams=None, perceptual_loss_fn=None, vgg_params=None, train=True,
            noise=None, bernoulli_u=None):
    """:100-186."""
    hparams = DEFAULT_HPARAMS if hparams is None else hparams
    reconstruction, compressed, selection, selection_mask, logvar, mean = model(
        video, mask, rngs, train=train, noise=noise, bernoulli_u=bernoulli_u)
    loss, aux = loss_terms(video, reconstruction, selection, selection_mask, logvar, mean, original_mask, hparams,
                           perceptual_loss_fn, vgg_params)
    aux.update(reconstruction=reconstruction, selection=selection, selection_mask=selection_mask)
    return loss, aux
