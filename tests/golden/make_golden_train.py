"""Fixture of the reference's TRAINING STEP over several updates (the optimizer row).

    python tests/golden/make_golden_train.py --shim          # this container: oracle/jaxshim (jax, flax.nnx, optax look-alikes)
    python tests/golden/make_golden_train.py                 # a box with real jax + flax + optax: writes jax_rltrain_*.npz

What runs, from /root/reference: train/rl_model.py (VideoVAE), and `per_sample_mean`, `magnify_negatives`, `loss_fn` and
`train_step` of train/rl_nonadversarial.py:59-198 exactly as written (taken from the file's AST, like make_golden_jax.py:
importing the module would pull in its dataloader / wandb / orbax).  The optimizer is built as the file's main() builds it
(:241-253): ``nnx.Optimizer(model, optax.chain(optax.clip_by_global_norm(1.0), optax.adam(learning_rate=schedule_fn)))``
with ``schedule_fn = optax.warmup_cosine_decay_schedule(0.0, peak, warmup, decay, peak / 10)`` -- a short warm-up and a
peak of 1e-3 here, so that a handful of steps walk through lr = 0 (the first update of a warm-up moves nothing), the
ramp and the first cosine steps, and the 1.0 global-norm clip is active on every step (the gradient norm of this model at
initialisation is > 1).

Written: per step the loss, its terms, the gradient global norm is NOT exported by the reference's step, so the fixture
holds what the step leaves behind -- after every update a 96-element probe + norm of every parameter, and at the end the
optimizer state as the reference checkpoints it (``nnx.state(optimizer)``, train/rl_nonadversarial.py:62-67): count, and
norm + probe of every first / second moment.  Inputs are by recipe (tests/golden/weight_recipe.py); the uniform draws
behind jax.random.bernoulli (rl) / the Gumbel gate (vae) are recorded per step.  Consumers: tests/test_jax_golden.py (oracle on CPU: oracle/rl_losses.py +
oracle/optim.py; the optimizer-state import of video_vae_b200/checkpoint.py)."""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_jax as G  # noqa: E402
import weight_recipe  # noqa: E402

SCHEDULE = dict(init_value=0.0, peak_value=1e-3, warmup_steps=3, decay_steps=40, end_value=1e-4)
STEPS = 6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--cfg", default="small", choices=["small", "hd64"])
    ap.add_argument("--steps", type=int, default=STEPS)
    ap.add_argument("--model", default="rl", choices=["rl", "vae"],
                    help="rl: train/rl_model.py + train_step of train/rl_nonadversarial.py (writes *_rltrain_*.npz); vae: "
                         "train/model.py + train_step AND eval_step of train/legacy/training_loop_adversarial.py:126-148 -- the "
                         "step bench.py times (writes *_vaetrain_*.npz)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--shim", action="store_true")
    args = ap.parse_args()
    if args.shim:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "jaxshim"))
    sys.path.insert(0, os.path.join(args.reference, "train"))
    import jax
    import jax.numpy as jnp
    import optax
    from flax import nnx
    if args.model == "rl":
        from rl_model import VideoVAE                              # the reference's train/rl_model.py
    else:
        from model import VideoVAE                                 # the reference's train/model.py
    shim = bool(getattr(jax, "IS_SHIM", False))
    assert shim == args.shim, "a jax look-alike is on sys.path without --shim (or --shim found the real jax first)"

    spec = G.CFGS[args.cfg]
    cfg, b, t = spec["cfg"], spec["batch"], spec["frames"]
    model = VideoVAE(*cfg, rngs=nnx.Rngs(2), dtype=jnp.float32, param_dtype=jnp.float32)
    state = nnx.state(model, nnx.Param)
    names = G.flatten_state(state)
    pure = state.to_pure_dict()

    def fill(d, path):
        for kk, vv in d.items():
            if isinstance(vv, dict):
                fill(vv, path + (str(kk),))
            else:
                name = ".".join(path + (str(kk),))
                assert name in names, name
                d[kk] = jnp.asarray(weight_recipe.param(name, vv.shape))      # final_conv included: non-zero by recipe
    fill(pure, ())
    if hasattr(nnx, "replace_by_pure_dict"):
        nnx.replace_by_pure_dict(state, pure)
        nnx.update(model, state)
    else:
        nnx.update(model, pure)

    # train/rl_nonadversarial.py:241-253
    schedule_fn = optax.warmup_cosine_decay_schedule(**SCHEDULE)
    optimizer_def = optax.chain(optax.clip_by_global_norm(1.0), optax.adam(learning_rate=schedule_fn))
    optimizer = nnx.Optimizer(model, optimizer_def)

    if args.model == "rl":
        ref_file = os.path.join(args.reference, "train", "rl_nonadversarial.py")
        train_step = G.reference_functions(ref_file, ("per_sample_mean", "magnify_negatives", "loss_fn", "train_step"), "train_step")
        hparams = G.RL_HPARAMS
    else:
        ref_file = os.path.join(args.reference, "train", "legacy", "training_loop_adversarial.py")
        fns = ("magnify_negatives", "loss_fn", "train_step", "eval_step")
        train_step = G.reference_functions(ref_file, fns, "train_step")
        eval_step = G.reference_functions(ref_file, fns, "eval_step")
        hparams = G.HPARAMS

    def perceptual(vgg_params, reconstruction, target):           # the stand-in of make_golden_jax.py (same reason)
        return jnp.mean(jnp.abs(reconstruction - target) ** 3, axis=(1, 2, 3, 4))

    hw = (cfg[0] // cfg[3]) * (cfg[1] // cfg[3])
    video = jnp.asarray(weight_recipe.clip((b, t, cfg[0], cfg[1], cfg[2])))
    original_mask = jnp.arange(t)[None, :] < jnp.asarray(spec["keep"])[:, None]
    out = {"cfg": np.asarray(cfg, np.int64), "dtype": np.asarray("float32"), "model": np.asarray(args.model + "_train"),
           "hparams": np.asarray(json.dumps(hparams)), "schedule": np.asarray(json.dumps(SCHEDULE)),
           "steps": np.asarray(args.steps, np.int64), "recipe": np.asarray(weight_recipe.RECIPE_ID),
           "video_shape": np.asarray(video.shape, np.int64), "mask": np.asarray(original_mask)}
    terms = ("MSE", "perceptual_loss", "selection_loss", "kl_loss", "kept_frame_density", "per_sample_MAE")
    for step in range(args.steps):
        with G.DrawRecorder(weight_recipe.normal) as rec:
            if args.model == "rl":
                loss, aux = train_step(model, optimizer, video, original_mask, hparams, hw, nnx.Rngs(100 + step), perceptual, None)
            else:
                loss, mse, sel_loss, kl, _recon, density = train_step(model, optimizer, video, original_mask, hparams, hw,
                                                                      nnx.Rngs(100 + step))
                aux = {"MSE": mse, "selection_loss": sel_loss, "kl_loss": kl, "kept_frame_density": density}
        assert len(rec.uniform) == 1 and len(rec.normal) == 1
        out[f"step{step}/bernoulli_u"] = rec.uniform[0].astype(np.float32)
        out[f"step{step}/noise_shape"] = np.asarray(rec.normal[0].shape, np.int64)
        out[f"step{step}/loss"] = np.asarray(loss, np.float32)
        for k in terms:
            if k in aux:
                out[f"step{step}/{k}"] = np.asarray(aux[k], np.float32)
        out[f"step{step}/lr"] = np.asarray(schedule_fn(step), np.float64)
        for name, v in G.flatten_state(nnx.state(model, nnx.Param)).items():
            v64 = v.astype(np.float64)
            out[f"step{step}/pnorm/{name}"] = np.asarray(np.sqrt((v64 * v64).sum()), np.float64)
            if step == args.steps - 1:
                out[f"final/pprobe/{name}"] = weight_recipe.grad_probe(v)
        print(f"step {step}: lr {float(schedule_fn(step)):.2e} loss {float(loss):.6f}", flush=True)

    if args.model == "vae":
        # eval_step (:139-148: train=False -- the latent is the mean, the gate has no noise) on the trained weights
        with G.DrawRecorder(weight_recipe.normal) as rec:
            loss, mse, sel_loss, kl, recon, density = eval_step(model, video, original_mask, hparams, hw, nnx.Rngs(7))
        assert len(rec.uniform) == 0 and len(rec.normal) == 0, "eval_step drew random numbers"
        for k, v in (("loss", loss), ("MSE", mse), ("selection_loss", sel_loss), ("kl_loss", kl), ("kept_frame_density", density)):
            out["eval/" + k] = np.asarray(v, np.float32)
        out["eval/reconstruction"] = np.ascontiguousarray(np.asarray(recon, np.float32)[:, :, ::3, ::3, :])

    # the optimizer half of the reference's checkpoint: nnx.state(optimizer) (train/rl_nonadversarial.py:62-67)
    opt = G.flatten_state(nnx.state(optimizer))
    n_mu = 0
    for name, v in opt.items():
        parts = name.split(".")
        if parts[0] == "model":
            continue
        if parts[-1] in ("count", "step") or "count" in parts:
            out["opt/" + name] = np.asarray(v).astype(np.int64)
        elif "mu" in parts or "nu" in parts:
            v64 = np.asarray(v, np.float64)
            out["opt_norm/" + name] = np.asarray(np.sqrt((v64 * v64).sum()), np.float64)
            out["opt_probe/" + name] = weight_recipe.grad_probe(np.asarray(v, np.float32))
            n_mu += "mu" in parts
    assert n_mu == len(names), (n_mu, len(names))
    out["generator"] = np.asarray(("reference files on oracle/jaxshim (CPU torch), jax " if shim else "reference files on jax ")
                                  + jax.__version__)
    path = args.out or os.path.join(HERE, f"{'refshim' if shim else 'jax'}_{args.model}train_{args.cfg}_float32.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
