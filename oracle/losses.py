"""Oracle restatement of the reference loss / step (CPU torch, TEST INFRASTRUCTURE ONLY).

Follows train/legacy/training_loop_adversarial.py:66-136 (magnify_negatives,
loss_fn, train_step mask plumbing) and the MAE term of
train/rl_nonadversarial.py:114-117.
"""
import torch

DEFAULT_HPARAMS = {  # training_loop_adversarial.py:47-48,52,54
    "gamma1": 0.05,
    "gamma2": 0.001,
    "max_compression_rate": 2,
    "magnify_negatives_rate": 100,
}


def magnify_negatives(x, rate):
    """training_loop_adversarial.py:66-68."""
    return torch.where(x < 0, x * rate, x)


def expand_mask(mask_bt, hw):
    """train_step, training_loop_adversarial.py:127-130: (b,t) -> ((b hw),1,1,t)."""
    b, t = mask_bt.shape
    return mask_bt.reshape(b, 1, 1, 1, t).expand(b, hw, 1, 1, t).reshape(b * hw, 1, 1, t)


def loss_terms(video, reconstruction, selection, logvar, mean, original_mask, hparams):
    """loss_fn body after the model call, training_loop_adversarial.py:94-124 (+ MAE)."""
    f32 = torch.promote_types(torch.float32, reconstruction.dtype)
    m = original_mask.to(f32)
    seq = torch.clamp(m.sum(dim=1, keepdim=True), min=1.0)                      # (b,1), fp32
    vm = original_mask.to(reconstruction.dtype)[:, :, None, None, None]
    err = (video.to(reconstruction.dtype) - reconstruction) * vm                # compute dtype
    sq = torch.square(err)
    frame_reduced = sq.to(f32).sum(dim=1, keepdim=True).to(sq.dtype).to(f32) / seq[:, :, None, None, None]
    mse = frame_reduced.mean()
    ab = torch.abs(err)
    mae = (ab.to(f32).sum(dim=1, keepdim=True).to(ab.dtype).to(f32) / seq[:, :, None, None, None]).mean()

    km = m[:, :, None, None]
    sel_sum = (selection.to(f32) * km).sum(dim=(1, 2, 3))[:, None]              # (b,1)
    density = sel_sum / seq
    diff = density - (1.0 / hparams["max_compression_rate"])
    selection_loss = torch.square(magnify_negatives(diff, hparams["magnify_negatives_rate"])).mean()

    lv, mu = logvar, mean
    kl_el = 0.5 * (torch.exp(lv) - 1 - lv + torch.square(mu))                   # compute dtype
    kl = (kl_el * original_mask.to(kl_el.dtype)[:, :, None, None]).to(f32) / seq[:, :, None, None]
    kl_loss = kl.mean()
    loss = mse + hparams["gamma1"] * selection_loss + hparams["gamma2"] * kl_loss
    return loss, {"MSE": mse, "MAE": mae, "selection_loss": selection_loss, "kl_loss": kl_loss,
                  "kept_frame_density": density.mean()}


def loss_fn(model, video, mask, original_mask, rngs, hparams=None, train=True, noise=None, gumbel_u=None):
    """training_loop_adversarial.py:90-124."""
    hparams = DEFAULT_HPARAMS if hparams is None else hparams
    reconstruction, compressed, selection, logvar, mean = model(video, mask, rngs, train=train,
                                                                noise=noise, gumbel_u=gumbel_u)
    loss, aux = loss_terms(video, reconstruction, selection, logvar, mean, original_mask, hparams)
    aux.update(reconstruction=reconstruction, compressed=compressed, selection=selection, logvar=logvar, mean=mean)
    return loss, aux


def eval_step(model, video, mask_bt, hparams, hw, rngs):
    """training_loop_adversarial.py:139-148."""
    with torch.no_grad():
        return loss_fn(model, video, expand_mask(mask_bt, hw), mask_bt, rngs, hparams, train=False)
