"""Encode-only latent extraction (BASELINE config 3).

The reference's ``data_prep/save_latents.py:183-206`` walks a clip list under ``torch.no_grad()``, encodes it in fixed
chunks in bf16, concatenates the chunk outputs on the host and saves ``{"latents": ...}`` with ``torch.save``.  (That
script drives a third-party image auto-encoder; this is the same loop over THIS repository's ``Encoder``,
``train/model.py:49-60`` in eval mode, where the Gumbel gate is deterministic.)
"""
import torch


@torch.no_grad()
def encode_latents(model, clips, mask=None, chunk=8):
    """clips [n, t, H, W, C] (any float dtype, host or device) -> dict(latents=mean [n,t,hw,Dl], log_variance=...,
    selection=[n,t] kept-frame gate), all on the host.  ``mask`` [n, t] bool (True = real frame), default all True."""
    enc = model.encoder
    dev = next(enc.parameters()).device
    n, t = clips.shape[:2]
    if mask is None:
        mask = torch.ones(n, t, dtype=torch.bool)
    means, logvars, sels = [], [], []
    from .rng import Rngs
    for i in range(0, n, chunk):
        x = clips[i:i + chunk].to(dev, non_blocking=True)
        m = mask[i:i + chunk].to(dev, non_blocking=True)
        mean, logvar, sel = enc(x, m[:, None, None, :], Rngs(0), train=False)
        means.append(mean.to("cpu"))
        logvars.append(logvar.to("cpu"))
        sels.append(sel.reshape(-1, t).to("cpu"))
    return {"latents": torch.cat(means), "log_variance": torch.cat(logvars), "selection": torch.cat(sels)}


def save_latents(model, clips, path, mask=None, chunk=8):
    """``torch.save`` of :func:`encode_latents` (the ``.pt`` layout of data_prep/save_latents.py:203)."""
    out = encode_latents(model, clips, mask=mask, chunk=chunk)
    torch.save(out, path)
    return out
