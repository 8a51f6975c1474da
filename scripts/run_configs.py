#!/usr/bin/env python
"""BASELINE.json configs[2..4] on one B200 (configs[1] is bench.py's headline; configs[0] is the CPU oracle case).

    python scripts/run_configs.py [cfg3] [cfg4] [cfg5]  -> one JSON line per config (also appended to gpurun_out/configs.jsonl)

cfg3: encode-only latent extraction, 32x256x256 clips, batch 32, bf16, eval mode, chunked like data_prep/save_latents.py:183-206
cfg4: cfg2 training step with prefix masks keeping 25/50/75/100 % of the 16 frames (train/dataloader.py:232-234 contract)
cfg5: long clip 64x512x512, batch 1, fwd+bwd bf16 (temporal L = 64, spatial L = 1024)
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import video_vae_b200 as V  # noqa: E402
from video_vae_b200.ddp import FlatParams  # noqa: E402

PROD = dict(patch_size=16, encoder_depth=9, decoder_depth=12, mlp_dim=1536, num_heads=8, qkv_features=512,
            max_temporal_len=64, spatial_compression_rate=8, unembedding_upsample_rate=4)


def build(size, dev):
    m = V.VideoVAE(size, size, 3, PROD["patch_size"], PROD["encoder_depth"], PROD["decoder_depth"], PROD["mlp_dim"],
                   PROD["num_heads"], PROD["qkv_features"], PROD["max_temporal_len"], PROD["spatial_compression_rate"],
                   PROD["unembedding_upsample_rate"], V.Rngs(2), dtype=torch.bfloat16, device=dev)
    with torch.no_grad():
        m.decoder.unet.final_conv.kernel.normal_(0.0, 0.02, generator=torch.Generator(device=dev).manual_seed(7))
    flat = FlatParams(m)
    flat.enable_bf16_shadow()
    return m, flat


def timed(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def cfg3(dev):
    m, _ = build(256, dev)
    B, T, chunk = 32, 32, 8
    g = torch.Generator().manual_seed(1234)
    video = torch.rand(B, T, 256, 256, 3, generator=g).to(torch.bfloat16).to(dev)
    mask = torch.ones(B, T, dtype=torch.bool, device=dev)

    def run():
        outs = []
        with torch.no_grad():
            for i in range(0, B, chunk):
                mean, logvar, sel = m.encoder(video[i:i + chunk], mask[i:i + chunk, None, None, :], V.Rngs(0), train=False)
                outs.append(mean)
        return torch.cat(outs)
    ms, lat = timed(run, 2, 5)
    tf = 38.9e12 / (ms * 1e-3) / 1e12
    return {"config": "cfg3 encode-only 32x256x256 batch 32 (chunks of 8), bf16, eval", "ms_per_batch": ms,
            "clips_per_s": B / (ms * 1e-3), "latent_shape": list(lat.shape), "model_tflops": tf,
            "finite": bool(torch.isfinite(lat.float()).all())}


def cfg4(dev):
    m, flat = build(256, dev)
    B, T = 8, 16
    g = torch.Generator().manual_seed(1234)
    video = torch.rand(B, T, 256, 256, 3, generator=g).to(torch.bfloat16).to(dev)
    res = {}
    hp = dict(V.DEFAULT_HPARAMS, gamma4=0.1)
    for keep in (4, 8, 12, 16):
        mask = (torch.arange(T)[None, :] < keep).expand(B, T).contiguous().to(dev)
        rngs = V.Rngs(3)

        def step():
            flat.zero_grad()
            loss, _ = V.loss_fn(m, video, mask[:, None, None, :], mask, rngs, hp, train=True)
            loss.backward()
            return loss
        ms, loss = timed(step, 2, 4)
        res[f"keep{keep}"] = {"ms_per_step": ms, "clips_per_s": B / (ms * 1e-3), "loss": loss.item()}
    return {"config": "cfg4 train step 16x256x256 batch 8 with prefix masks (frames kept of 16)", **res}


def cfg5(dev):
    m, flat = build(512, dev)
    B, T = 1, 64
    g = torch.Generator().manual_seed(1234)
    video = torch.rand(B, T, 512, 512, 3, generator=g).to(torch.bfloat16).to(dev)
    mask = torch.ones(B, T, dtype=torch.bool, device=dev)
    hp = dict(V.DEFAULT_HPARAMS, gamma4=0.1)
    rngs = V.Rngs(3)

    def step():
        flat.zero_grad()
        loss, _ = V.loss_fn(m, video, mask[:, None, None, :], mask, rngs, hp, train=True)
        loss.backward()
        return loss
    ms, loss = timed(step, 1, 3)
    return {"config": "cfg5 long clip 64x512x512 batch 1 fwd+bwd bf16", "ms_per_step": ms, "clips_per_s": B / (ms * 1e-3),
            "model_tflops": 88.5e12 / (ms * 1e-3) / 1e12, "loss": loss.item(),
            "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}


def rl(dev):
    """Production RL step of train/rl_nonadversarial.py (:36-57, 100-198): rl_model, batch 2 x 32 frames (duplicated
    to 4 inside the model), RL loss with and without the VGG perceptual term (random VGG weights)."""
    from video_vae_b200.perceptual import get_adversarial_perceptual_loss_fn, load_vgg
    from video_vae_b200.rl_losses import DEFAULT_HPARAMS as RL_HP, loss_fn as rl_loss_fn
    from video_vae_b200.rl_model import VideoVAE as RLVAE
    m = RLVAE(256, 256, 3, PROD["patch_size"], PROD["encoder_depth"], PROD["decoder_depth"], PROD["mlp_dim"],
              PROD["num_heads"], PROD["qkv_features"], PROD["max_temporal_len"], PROD["spatial_compression_rate"],
              PROD["unembedding_upsample_rate"], V.Rngs(2), dtype=torch.bfloat16, device=dev)
    with torch.no_grad():
        m.decoder.unet.final_conv.kernel.normal_(0.0, 0.02, generator=torch.Generator(device=dev).manual_seed(7))
    flat = FlatParams(m)
    flat.enable_bf16_shadow()
    vgg, vgg_params = load_vgg(V.Rngs(5), device=dev)
    pfn = get_adversarial_perceptual_loss_fn(vgg)
    B, T = 2, 32
    g = torch.Generator().manual_seed(1234)
    video = torch.rand(B, T, 256, 256, 3, generator=g).to(torch.bfloat16).to(dev)
    mask = torch.ones(B, T, dtype=torch.bool, device=dev)
    res = {}
    # graphs first: a capture after eager backward passes trips over their cross-stream dependencies
    from video_vae_b200.graph import GraphedRLTrainStep
    for name, fn, hp in (("graph_no_perceptual", None, dict(RL_HP, gamma3=0.0)), ("graph_vgg_perceptual", pfn, dict(RL_HP))):
        gs = GraphedRLTrainStep(m, flat, video, mask, hp, perceptual_loss_fn=fn, vgg_params=vgg_params)
        rngs = V.Rngs(3)
        ms, loss = timed(lambda: gs(video, mask, rngs), 2, 6)
        res[name] = {"ms_per_step": ms, "clips_per_s": B / (ms * 1e-3), "loss": loss.item(),
                     "grad_finite": bool(torch.isfinite(flat.grad).all())}
        del gs
    for name, fn, hp in (("no_perceptual", None, dict(RL_HP, gamma3=0.0)), ("vgg_perceptual", pfn, dict(RL_HP))):
        rngs = V.Rngs(3)

        def step():
            flat.zero_grad()
            loss, aux = rl_loss_fn(m, video, mask[:, None, None, :], mask, rngs, hp, fn, vgg_params, train=True)
            loss.backward()
            return loss, aux
        ms, (loss, aux) = timed(step, 2, 4)
        res[name] = {"ms_per_step": ms, "clips_per_s": B / (ms * 1e-3), "loss": loss.item(),
                     "perceptual_loss": float(aux["perceptual_loss"]), "grad_finite": bool(torch.isfinite(flat.grad).all())}
    return {"config": "rl_nonadversarial production step: rl_model, batch 2 x 32 frames x 256^2 (decoder batch 4), bf16, "
                      "fwd+loss+bwd, eager and CUDA-graph", **res, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as f:
        for name in which:
            t0 = time.time()
            try:
                r = {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "rl": rl}[name](dev)
            except Exception as e:  # noqa: BLE001
                r = {"config": name, "error": repr(e)[:500]}
            r["wall_s"] = round(time.time() - t0, 1)
            print(json.dumps(r), flush=True)
            f.write(json.dumps(r) + "\n")
            torch.cuda.empty_cache()
