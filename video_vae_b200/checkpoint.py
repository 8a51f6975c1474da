"""Weight import / export against the reference's Flax parameter tree.

The reference checkpoints ``{"model": nnx.state(model), "optimizer": nnx.state(optimizer)}`` with orbax
(train/rl_nonadversarial.py:62-67, train/model_loader.py:35-42).  ``nnx.state`` is a nested mapping keyed by attribute
names and list indices whose leaves are arrays (older Flax) or ``{"value": array}`` variable states (Flax >= 0.10).
This package keeps the reference's attribute names AND Flax array layouts (Linear ``(in, out)``, Conv
``(kt, kh, kw, Cin, Cout)``, ConvTranspose kernels as stored), so a tree converts with a flat rename: no transposes.

Nothing here touches the GPU.  orbax / jax are not installable in the build image; ``load_checkpoint`` uses orbax when
it is importable and otherwise reads the ``.npz`` transport format written by ``save_npz`` (a reference maintainer dumps
``nnx.state(model)`` with ``save_npz(path, jax.tree.map(np.asarray, nnx.state(model).to_pure_dict()))``).
"""
import os

import numpy as np
import torch

_VALUE_KEYS = ("value", "raw_value")


def flatten_tree(tree, prefix=""):
    """Nested mapping / sequence -> {dotted.path: ndarray}.  ``{"value": x}`` variable wrappers are unwrapped."""
    out = {}
    if isinstance(tree, dict):
        keys = list(tree.keys())
        if len(keys) == 1 and keys[0] in _VALUE_KEYS and not isinstance(tree[keys[0]], (dict, list, tuple)):
            out[prefix[:-1]] = np.asarray(tree[keys[0]])
            return out
        for k in keys:
            out.update(flatten_tree(tree[k], f"{prefix}{k}."))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out.update(flatten_tree(v, f"{prefix}{i}."))
    elif tree is not None:
        out[prefix[:-1]] = np.asarray(tree)
    return out


def _is_rng_state(name):
    parts = name.split(".")
    return any(p in ("rngs", "rng", "dropout") for p in parts)


_ROPE_TABLES = ("cos_cached", "sin_cached")


def _is_rope_table(name):
    """``RotaryEmbedding.{cos,sin}_cached`` are ``nnx.Variable``s in the reference (train/layers.py:101-102), so an
    unfiltered ``nnx.state(model)`` carries them (``...Attention.ROPE.cos_cached``, shape [1, max_len, 1, hd]).  Here
    they are non-persistent buffers rebuilt from (head_dim, max_len, base): not parameters, never loaded."""
    parts = name.split(".")
    return len(parts) >= 2 and parts[-1] in _ROPE_TABLES and parts[-2] == "ROPE"


def rope_tables_match(tree, model, rtol=1e-5):
    """Check the checkpoint's RoPE tables against the ones this model recomputes (same formula: layers.py:87-102).
    Returns the number of tables compared; raises ValueError on the first mismatch."""
    flat = {k: v for k, v in flatten_tree(tree).items() if _is_rope_table(k)}
    mods = dict(model.named_modules())
    n = 0
    for k, v in flat.items():
        owner, leaf = k.rsplit(".", 1)
        rope = mods.get(owner)
        if rope is None:
            raise ValueError(f"{k}: no RotaryEmbedding at {owner}")
        mine = getattr(rope, leaf).detach().float().cpu().numpy().reshape(-1)
        theirs = np.asarray(v, dtype=np.float32).reshape(-1)
        if mine.shape != theirs.shape or not np.allclose(mine, theirs, rtol=rtol, atol=1e-6):
            raise ValueError(f"{k}: checkpoint RoPE table differs from the recomputed one")
        n += 1
    return n


def state_dict_from_flax(tree, model, strict=True):
    """Flax parameter tree (nested or already flat) -> ``state_dict`` for ``model`` (shape- and name-checked)."""
    flat = flatten_tree(tree)
    flat = {k: v for k, v in flat.items() if not _is_rng_state(k) and not _is_rope_table(k)}
    want = model.state_dict()
    missing = [k for k in want if k not in flat]
    unexpected = [k for k in flat if k not in want]
    if strict and (missing or unexpected):
        raise KeyError(f"checkpoint/model mismatch: missing {missing[:5]} (+{max(0, len(missing) - 5)}), "
                       f"unexpected {unexpected[:5]} (+{max(0, len(unexpected) - 5)})")
    sd = {}
    for k, ref in want.items():
        if k not in flat:
            continue
        a = np.asarray(flat[k])
        if a.dtype.kind == "V" or str(a.dtype) == "bfloat16":          # ml_dtypes bfloat16 -> fp32 master copy
            a = a.astype(np.float32)
        if tuple(a.shape) != tuple(ref.shape):
            raise ValueError(f"{k}: checkpoint shape {tuple(a.shape)} != model shape {tuple(ref.shape)}")
        sd[k] = torch.from_numpy(np.ascontiguousarray(a)).to(ref.dtype)
    return sd


def load_flax_tree(model, tree, strict=True, flat=None):
    """Load a Flax tree into ``model``; ``flat`` (a ``ddp.FlatParams``) gets its compute-dtype shadow refreshed."""
    sd = state_dict_from_flax(tree, model, strict=strict)
    with torch.no_grad():
        own = model.state_dict()
        for k, v in sd.items():
            own[k].copy_(v)
    if flat is not None and getattr(flat, "shadow", None) is not None:
        flat.refresh_shadow()
    else:
        from . import functional as F_
        F_.invalidate_shadows()
    return sorted(sd)


def to_flax_tree(model, wrap_value=False):
    """``model`` -> nested dict shaped like ``nnx.state(model)`` (list indices become int keys)."""
    tree = {}
    for name, t in model.state_dict().items():
        node = tree
        parts = name.split(".")
        for p in parts[:-1]:
            node = node.setdefault(int(p) if p.isdigit() else p, {})
        a = t.detach().to("cpu", torch.float32).numpy()
        node[parts[-1]] = {"value": a} if wrap_value else a
    return tree


def load_adam_moments(adam, model, mu_tree, nu_tree, count):
    """optax ``ScaleByAdamState(count, mu, nu)`` (trees shaped like the parameters) -> ``ddp.FlatAdam`` buffers."""
    fp = adam.flat
    names = {id(p): n for n, p in model.named_parameters()}
    mu, nu = flatten_tree(mu_tree), flatten_tree(nu_tree)
    with torch.no_grad():
        for i, p in enumerate(fp.params):
            n = names[id(p)]
            s, e = fp.offsets[i], fp.offsets[i] + p.numel()
            for buf, src in ((adam.m, mu), (adam.v, nu)):
                a = np.asarray(src[n], dtype=np.float32)
                if a.size != p.numel():
                    raise ValueError(f"{n}: moment has {a.size} elements, parameter {p.numel()}")
                buf[s:e].copy_(torch.from_numpy(a.reshape(-1)))
    adam.t = int(count)


def _find_adam_state(opt_tree):
    """Locate optax's ``ScaleByAdamState(count, mu, nu)`` inside ``nnx.state(optimizer)`` /
    ``optimizer.opt_state``: the chain ``optax.chain(clip_by_global_norm(1.0), adam(schedule))`` of
    train/rl_nonadversarial.py:241-253 nests it as ``opt_state -> 1 -> 0`` (clip's EmptyState is element 0; adam is
    itself ``chain(scale_by_adam, scale_by_learning_rate)``), but the position is found by its field names so other
    chains work too.  Returns (mu flat dict, nu flat dict, count or None)."""
    flat = flatten_tree(opt_tree)
    mu, nu, counts = {}, {}, {}
    for k, v in flat.items():
        parts = k.split(".")
        for field, dst in (("mu", mu), ("nu", nu)):
            if field in parts:
                i = parts.index(field)
                dst[(".".join(parts[:i]), ".".join(parts[i + 1:]))] = v
        if parts[-1] == "count":
            counts[".".join(parts[:-1])] = v
    if not mu or not nu:
        raise KeyError("no ScaleByAdamState (mu / nu subtrees) found in the optimizer state")
    prefixes = {pfx for pfx, _ in mu}
    if len(prefixes) != 1 or prefixes != {pfx for pfx, _ in nu}:
        raise KeyError(f"ambiguous Adam state: mu/nu found under {sorted(prefixes)}")
    pfx = prefixes.pop()
    count = counts.get(pfx)
    return ({n: v for (_, n), v in mu.items()}, {n: v for (_, n), v in nu.items()},
            None if count is None else int(np.asarray(count).reshape(-1)[0]))


def load_optimizer_state(adam, model, opt_tree, count=None):
    """The optimizer half of the reference's ``{"model", "optimizer"}`` checkpoint (train/rl_nonadversarial.py:62-67:
    ``nnx.state(optimizer)``) -> ``ddp.FlatAdam``: first / second moments and the step count.  RoPE tables and RNG
    leaves inside mu / nu (zero moments of non-trainable variables) are ignored."""
    mu, nu, found = _find_adam_state(opt_tree)
    keep = lambda d: {k: v for k, v in d.items() if not _is_rope_table(k) and not _is_rng_state(k)}   # noqa: E731
    if count is None:
        count = found
    if count is None:
        step = flatten_tree(opt_tree).get("step")
        count = 0 if step is None else int(np.asarray(step).reshape(-1)[0])
    load_adam_moments(adam, model, keep(mu), keep(nu), count)
    return count


def save_npz(path, tree):
    flat = flatten_tree(tree)
    np.savez(path, **{k: np.asarray(v, dtype=np.float32) if np.asarray(v).dtype.kind == "f" else np.asarray(v)
                      for k, v in flat.items()})
    return sorted(flat)


def load_npz(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def save_checkpoint(model, path):
    """Export in the transport format (the reference re-imports it with nnx.update on the unflattened tree)."""
    return save_npz(path, to_flax_tree(model))


def load_checkpoint(model, path, strict=True, flat=None, adam=None):
    """train/model_loader.py:35-42: an orbax directory (needs orbax) or a ``save_npz`` file.  The model half is always
    loaded; with ``adam`` (a ``ddp.FlatAdam``) the optimizer half (``nnx.state(optimizer)``) is mapped onto it too."""
    if os.path.isdir(path):
        try:
            import orbax.checkpoint as ocp                                          # noqa: PLC0415
        except ImportError as e:
            raise ImportError("reading an orbax checkpoint directory needs orbax-checkpoint (not in this image); "
                              "convert it to .npz on a JAX box with checkpoint.save_npz") from e
        restored = ocp.StandardCheckpointer().restore(os.path.abspath(path))
        tree = restored["model"] if isinstance(restored, dict) and "model" in restored else restored
        if adam is not None and isinstance(restored, dict) and "optimizer" in restored:
            load_optimizer_state(adam, model, restored["optimizer"])
    else:
        tree = load_npz(path)
        if any(k.startswith("model.") for k in tree):          # a flattened {"model": ..., "optimizer": ...} dump
            opt = {k[len("optimizer."):]: v for k, v in tree.items() if k.startswith("optimizer.")}
            tree = {k[len("model."):]: v for k, v in tree.items() if k.startswith("model.")}
            if adam is not None and opt:
                load_optimizer_state(adam, model, _unflatten(opt))
    return load_flax_tree(model, tree, strict=strict, flat=flat)


def _unflatten(flat):
    tree = {}
    for k, v in flat.items():
        node = tree
        parts = k.split(".")
        for p_ in parts[:-1]:
            node = node.setdefault(p_, {})
        node[parts[-1]] = v
    return tree
