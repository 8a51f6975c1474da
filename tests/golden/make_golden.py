#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ from the CPU oracle (fp32).

    python tests/golden/make_golden.py

The reference (JAX/Flax) cannot be imported in this container, so these vectors pin the ORACLE (oracle/*.py, the CPU
restatement of train/layers.py, train/model.py, train/unet.py and the loss of
train/legacy/training_loop_adversarial.py:90-124), not the reference itself ("parity unpinned").
They make the oracle's behaviour a committed artefact: tests/test_oracle.py re-derives them on CPU and
tests/test_parity_gpu.py checks the CUDA path against them at the fp32 tolerance of north_star (rel 1e-4).
Inputs and weights are regenerated from seeds by `build_case`; only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = (64, 64, 3, 16, 2, 2, 256, 4, 128, 32, 8, 4)   # the reference's CPU test config (claude_distributed/test_rl_model.py)
B, T = 2, 6


def build_case():
    """Seeded oracle model + inputs shared by the generator and the tests."""
    from oracle import Rngs
    from oracle.model import VideoVAE
    torch.manual_seed(0)
    model = VideoVAE(*CFG, Rngs(2))
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():   # the reference zero-inits final_conv (train/unet.py:144-153): randomise it so the UNet matters
        model.decoder.unet.final_conv.kernel.copy_(torch.randn(model.decoder.unet.final_conv.kernel.shape, generator=g) * 0.05)
    hw = (CFG[0] // CFG[3]) * (CFG[1] // CFG[3])
    video = torch.rand(B, T, CFG[0], CFG[1], 3, generator=g)
    mask = torch.tensor([[True] * T, [True] * (T - 2) + [False] * 2])
    noise = torch.randn(B, T, hw, CFG[3] * CFG[3] * 3 // CFG[10], generator=g)
    gumbel_u = torch.rand(B, T, 1, generator=g)
    return model, video, mask, noise, gumbel_u, hw


def run_oracle():
    from oracle import Rngs
    from oracle.losses import DEFAULT_HPARAMS, expand_mask, loss_fn
    model, video, mask, noise, gumbel_u, hw = build_case()
    loss, aux = loss_fn(model, video, expand_mask(mask, hw), mask, Rngs(0), DEFAULT_HPARAMS, noise=noise, gumbel_u=gumbel_u)
    loss.backward()
    out = {"loss": loss.detach().numpy(), "MSE": aux["MSE"].detach().numpy(), "MAE": aux["MAE"].detach().numpy(),
           "kl_loss": aux["kl_loss"].detach().numpy(), "selection_loss": aux["selection_loss"].detach().numpy(),
           "selection": aux["selection"].detach().reshape(B, T).numpy(),
           "mean_slice": aux["mean"].detach()[:, :, ::5, ::7].numpy(),
           "logvar_slice": aux["logvar"].detach()[:, :, ::5, ::7].numpy(),
           "compressed_slice": aux["compressed"].detach()[:, :, ::5, ::7].numpy(),
           "recon_slice": aux["reconstruction"].detach()[:, :, ::9, ::11, :].numpy()}
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    out["grad_qkv_slice"] = model.encoder.layers[0].TemporalAttention.qkv_projection.kernel.grad[::16, ::32].numpy()
    out["grad_conv_slice"] = model.decoder.unet.encoders[0].conv1.conv.kernel.grad[:, :, :, ::4, ::4].numpy()
    return out


def run_attention_kat():
    """train/attention_mask_tests.py shapes: q,k,v [17,15,19,13], the last 5 keys masked."""
    from oracle.nn import dot_product_attention
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(17, 15, 19, 13, generator=g) for _ in range(3))
    mask = torch.ones(17, 19, 15, 15, dtype=torch.bool)
    mask[..., 10:] = False
    o = dot_product_attention(q, k, v, mask)
    return {"attn_out_slice": o[::4, :, ::6, :].numpy(), "attn_out_sum": np.float64(o.double().sum())}


def build_rl_case():
    """Seeded oracle RL model, VGG feature extractor and inputs (rows next to the path)."""
    from oracle import Rngs
    from oracle.perceptual import VGG16Features
    from oracle.rl_model import VideoVAE
    model = VideoVAE(*CFG, Rngs(2))
    g = torch.Generator().manual_seed(23)
    with torch.no_grad():
        model.decoder.unet.final_conv.kernel.copy_(torch.randn(model.decoder.unet.final_conv.kernel.shape, generator=g) * 0.05)
    vgg = VGG16Features(Rngs(4))
    hw = (CFG[0] // CFG[3]) * (CFG[1] // CFG[3])
    b, t = 2, 4
    video = torch.rand(b, t, CFG[0], CFG[1], 3, generator=g)
    mask = torch.tensor([[True] * t, [True] * (t - 1) + [False]])
    noise = torch.randn(b, t, hw, CFG[3] * CFG[3] * 3 // CFG[10], generator=g)
    bernoulli_u = torch.rand(2 * b, t, 1, 1, generator=g)
    bernoulli_u[0::2, 0] = 0.0          # twins differ on frame 0, so the pairwise disadvantages are well conditioned
    bernoulli_u[1::2, 0] = 1.0
    return model, vgg, video, mask, noise, bernoulli_u


RL_HP = {"gamma1": 0.2, "gamma2": 0.001, "gamma3": 0.1, "gamma4": 0.05, "max_compression_rate": 2,
         "magnify_negatives_rate": 100, "rl_loss_weight": 0.5}


def run_rl_oracle():
    from oracle import Rngs
    from oracle.optim import ClipAdam
    from oracle.perceptual import get_adversarial_perceptual_loss_fn
    from oracle.rl_losses import loss_fn
    model, vgg, video, mask, noise, bu = build_rl_case()
    pfn = get_adversarial_perceptual_loss_fn(vgg)
    loss, aux = loss_fn(model, video, mask[:, None, None, :], mask, Rngs(0), RL_HP, pfn, None, noise=noise, bernoulli_u=bu)
    loss.backward()
    out = {"rl_loss_total": loss.detach().numpy(), "rl_per_sample_loss": aux["per_sample_loss"].detach().numpy(),
           "rl_selection_mask": aux["selection_mask"].detach().reshape(4, -1).numpy(),
           "rl_selection": aux["selection"].detach().reshape(4, -1).numpy(),
           "rl_recon_slice": aux["reconstruction"].detach()[:, :, ::9, ::11, :].numpy()}
    for k in ("MSE", "perceptual_loss", "selection_loss", "kl_loss", "kept_frame_density", "mean_trajectory_prob",
              "per_sample_MAE"):
        out["rl_" + k] = aux[k].detach().numpy()
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    out["rl_grad_names"] = np.array(names)
    out["rl_grad_norms"] = np.array(norms, dtype=np.float64)
    out["rl_grad_sel2"] = model.encoder.selection_layer2.kernel.grad.numpy()
    # one clip+Adam step on those gradients (optax.chain(clip_by_global_norm(1.0), adam(1e-3)))
    params = [p for p in model.parameters()]
    opt = ClipAdam(params, lr=1e-3, clip=1.0)
    gn = opt.step([p.grad if p.grad is not None else torch.zeros_like(p) for p in params])
    out["rl_grad_global_norm"] = np.float64(gn)
    out["rl_fill_token_after_step"] = model.fill_token.detach().reshape(-1)[:16].numpy()
    out["rl_qkv_after_step_slice"] = model.encoder.layers[0].TemporalAttention.qkv_projection.kernel.detach()[::16, ::32].numpy()
    return out


# ---------------------------------------------------------------------------------------------- production depth
def prod_cfg(size):
    """Production hyper-parameters (train/rl_nonadversarial.py:234-236) at a square clip size."""
    return (size, size, 3, 16, 9, 12, 1536, 8, 512, 64, 8, 4)


PROD_KEEP_FRAMES = (1, 0, 1, 1, 0, 1, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1)


def prod_inputs(size, keep, seed=11):
    """Seeded 16-frame clip, prefix mask keeping `keep` frames (train/dataloader.py:232-234), reparameterisation noise
    and DECISIVE Gumbel draws: u = sigmoid(+-6), so the logistic noise of train/layers.py:246-248 is +-6 and the gate
    sigmoid(logit + noise) rounds the same way in fp32 and bf16 (|logit| << 6 at initialisation)."""
    g = torch.Generator().manual_seed(seed)
    hw = (size // 16) ** 2
    video = torch.rand(1, 16, size, size, 3, generator=g)
    mask = torch.zeros(1, 16, dtype=torch.bool)
    mask[:, :keep] = True
    noise = torch.randn(1, 16, hw, 96, generator=g)
    keep_frame = torch.tensor(PROD_KEEP_FRAMES, dtype=torch.float32)
    u = torch.sigmoid((keep_frame * 2 - 1) * 6.0).reshape(1, 16, 1)
    return video, mask, noise, u, hw, keep_frame


def build_prod_model(size):
    from oracle import Rngs
    from oracle.model import VideoVAE
    o = VideoVAE(*prod_cfg(size), Rngs(2))
    with torch.no_grad():   # final_conv is zero-initialised in the reference (train/unet.py:144-153): make the U-Net matter
        k = o.decoder.unet.final_conv.kernel
        k.copy_(torch.randn(k.shape, generator=torch.Generator().manual_seed(5)) * 0.05)
    return o


def run_prod_step(o, size, keep):
    """Oracle forward + loss + backward at production depth; returns (loss, aux) with gradients left in o's .grad."""
    from oracle import Rngs
    from oracle.losses import DEFAULT_HPARAMS, expand_mask, loss_fn
    video, mask, noise, u, hw, _ = prod_inputs(size, keep)
    for p in o.parameters():
        p.grad = None
    loss, aux = loss_fn(o, video, expand_mask(mask, hw), mask, Rngs(0), DEFAULT_HPARAMS, noise=noise, gumbel_u=u)
    loss.backward()
    return loss, aux


def run_prod_oracle(size=128, keep=12):
    """BASELINE configs[0] (one 16x128x128 clip, fp32) at production depth, 12 of 16 frames kept."""
    o = build_prod_model(size)
    loss, aux = run_prod_step(o, size, keep)
    out = {"loss": loss.detach().numpy(), "MSE": aux["MSE"].detach().numpy(), "MAE": aux["MAE"].detach().numpy(),
           "kl_loss": aux["kl_loss"].detach().numpy(), "selection_loss": aux["selection_loss"].detach().numpy(),
           "selection": aux["selection"].detach().reshape(1, 16).numpy(),
           "mean_slice": aux["mean"].detach()[:, :, ::5, ::7].numpy(),
           "logvar_slice": aux["logvar"].detach()[:, :, ::5, ::7].numpy(),
           "recon_slice": aux["reconstruction"].detach()[:, :, ::9, ::11, :].numpy()}
    names, norms = [], []
    for n, p in o.named_parameters():
        names.append(n)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    out["grad_qkv_slice"] = o.encoder.layers[8].SpatialAttention.qkv_projection.kernel.grad[::16, ::32].numpy()
    out["grad_mlp_slice"] = o.decoder.layers[0].TemporalMLP.linear1.kernel.grad[::16, ::32].numpy()
    return out


def main():
    out = run_oracle()
    out.update(run_attention_kat())
    path = os.path.join(HERE, "videovae_cfg64_fp32.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; loss =", float(out["loss"]))
    out = run_rl_oracle()
    path = os.path.join(HERE, "rl_step_cfg64_fp32.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; loss =", float(out["rl_loss_total"]))
    out = run_prod_oracle()
    path = os.path.join(HERE, "videovae_prod128_fp32.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; loss =", float(out["loss"]))


if __name__ == "__main__":
    main()
