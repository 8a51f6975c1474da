"""B200-native loss / step of the reference's training loop.

``loss_fn`` mirrors train/legacy/training_loop_adversarial.py:90-124 (signature, loss terms, returned aux), with the
optional MAE term of train/rl_nonadversarial.py:114-117 enabled by ``hparams["gamma4"]``; ``train_step`` mirrors
:126-136 (mask plumbing + value_and_grad), leaving the optimizer update to the caller; ``eval_step`` mirrors :139-148.
"""
import torch

from . import functional as F_

DEFAULT_HPARAMS = {  # training_loop_adversarial.py:47-48,52,54
    "gamma1": 0.05,
    "gamma2": 0.001,
    "max_compression_rate": 2,
    "magnify_negatives_rate": 100,
}


def expand_mask(mask_bt, hw):
    """(b,t) -> ((b hw),1,1,t) as train_step builds it (training_loop_adversarial.py:127-130).  Provided for drop-in
    callers; the kernels index the (b,t) mask directly, so passing ``mask_bt[:, None, None, :]`` avoids the copy."""
    b, t = mask_bt.shape
    return mask_bt.reshape(b, 1, 1, 1, t).expand(b, hw, 1, 1, t).reshape(b * hw, 1, 1, t)


def loss_fn(model, video, mask, original_mask, rngs, hparams=None, train=True, noise=None, gumbel_u=None):
    hparams = DEFAULT_HPARAMS if hparams is None else hparams
    reconstruction, compressed, selection, logvar, mean = model(video, mask, rngs, train=train, noise=noise,
                                                                gumbel_u=gumbel_u)
    loss, mse, mae, sel_loss, kl, density = F_.VaeLossFn.apply(video, reconstruction, selection, logvar, mean,
                                                               original_mask, hparams)
    aux = {"MSE": mse, "MAE": mae, "selection_loss": sel_loss, "kl_loss": kl, "kept_frame_density": density,
           "reconstruction": reconstruction, "compressed": compressed, "selection": selection, "logvar": logvar,
           "mean": mean}
    return loss, aux


def train_step(model, video, mask_bt, hparams, rngs, noise=None, gumbel_u=None):
    """Forward + backward of one batch; gradients are left accumulated in ``p.grad`` (fp32)."""
    loss, aux = loss_fn(model, video, mask_bt[:, None, None, :], mask_bt, rngs, hparams, train=True, noise=noise,
                        gumbel_u=gumbel_u)
    loss.backward()
    return loss, aux


def eval_step(model, video, mask_bt, hparams, rngs):
    """training_loop_adversarial.py:139-148: the loss with ``train=False`` -- the latent is the mean (model.py:113-125)
    and the frame gate is the deterministic threshold ``round(sigmoid(logit))`` (layers.py:250-252); no draws, no gradients."""
    with torch.no_grad():
        return loss_fn(model, video, mask_bt[:, None, None, :], mask_bt, rngs, hparams, train=False)
