"""`optax` stand-in on CPU torch -- the slice the reference's training code uses.  TEST INFRASTRUCTURE ONLY.

optax (pinned 0.2.6, claude_distributed/requirements.txt) is not vendored by the reference and cannot be installed
offline.  The reference builds ``optax.chain(optax.clip_by_global_norm(1.0), optax.adam(learning_rate=schedule_fn))`` with
``optax.warmup_cosine_decay_schedule`` (train/rl_nonadversarial.py:241-253, claude_distributed/distributed_train.py:364-374)
and ``optax.adam(LR)`` + ``optax.apply_updates`` (claude_distributed/distributed_rl_model.py:65,91).  Restated here from
optax's published algorithms, independently of oracle/optim.py (which fuses the chain into one loop): this file keeps
optax's own structure -- GradientTransformation(init, update) pairs over pytrees, chain state = tuple of member states,
adam = chain(scale_by_adam, scale_by_learning_rate) -- so that ``nnx.state(optimizer)`` has the nesting a real checkpoint
has (opt_state -> 1 -> 0 -> {count, mu, nu}) and the two restatements check each other (tests/test_jaxshim_cpu.py).
Trees are the shim's nnx.State or nested dicts of arrays."""
import math
from collections import namedtuple

import torch

GradientTransformation = namedtuple("GradientTransformation", ["init", "update"])
EmptyState = namedtuple("EmptyState", [])
ScaleByAdamState = namedtuple("ScaleByAdamState", ["count", "mu", "nu"])
ScaleByScheduleState = namedtuple("ScaleByScheduleState", ["count"])


def _flat(tree):
    if hasattr(tree, "flat_state"):
        return tree.flat_state()
    out = {}

    def walk(d, path):
        for k, v in d.items():
            walk(v, path + (k,)) if isinstance(v, dict) else out.__setitem__(path + (k,), v)
    walk(tree, ())
    return out


def _like(tree, flat):
    if hasattr(tree, "flat_state"):
        return type(tree)(flat)
    root = {}
    for path, v in flat.items():
        d = root
        for k in path[:-1]:
            d = d.setdefault(k, {})
        d[path[-1]] = v
    return root


def _map(fn, tree, *rest):
    flats = [_flat(r) for r in rest]
    return _like(tree, {p: fn(v, *[f[p] for f in flats]) for p, v in _flat(tree).items()})


def global_norm(tree):
    return torch.sqrt(sum((v.detach().float() ** 2).sum() for v in _flat(tree).values()))


def clip_by_global_norm(max_norm):
    """updates <- updates if ||updates|| < max_norm else updates / ||updates|| * max_norm (norm over ALL leaves)."""
    def update(updates, state, params=None):
        g_norm = global_norm(updates)
        if bool(g_norm < max_norm):
            return updates, state
        return _map(lambda u: u / g_norm.to(u.dtype) * max_norm, updates), state
    return GradientTransformation(lambda params: EmptyState(), update)


def scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        zeros = lambda p: torch.zeros_like(torch.as_tensor(p))                                 # noqa: E731
        return ScaleByAdamState(torch.zeros((), dtype=torch.int32), _map(zeros, params), _map(zeros, params))

    def update(updates, state, params=None):
        mu = _map(lambda g, m: (1 - b1) * g + b1 * m, updates, state.mu)
        nu = _map(lambda g, n: (1 - b2) * (g * g) + b2 * n, updates, state.nu)
        count = state.count + 1
        t = int(count)
        mu_hat = _map(lambda m: m / (1 - b1 ** t), mu)
        nu_hat = _map(lambda n: n / (1 - b2 ** t), nu)
        out = _map(lambda m, n: m / (torch.sqrt(n + eps_root) + eps), mu_hat, nu_hat)
        return out, ScaleByAdamState(count, mu, nu)
    return GradientTransformation(init, update)


def scale_by_learning_rate(learning_rate):
    """Multiply by -learning_rate; a schedule is evaluated at the number of updates made BEFORE this one (first: 0)."""
    if callable(learning_rate):
        def update(updates, state, params=None):
            lr = float(learning_rate(int(state.count)))
            return _map(lambda u: -lr * u, updates), ScaleByScheduleState(state.count + 1)
        return GradientTransformation(lambda params: ScaleByScheduleState(torch.zeros((), dtype=torch.int32)), update)
    return GradientTransformation(lambda params: EmptyState(),
                                  lambda updates, state, params=None: (_map(lambda u: -learning_rate * u, updates), state))


def chain(*transforms):
    def init(params):
        return tuple(t.init(params) for t in transforms)

    def update(updates, state, params=None):
        new = []
        for t, s in zip(transforms, state):
            updates, s = t.update(updates, s, params)
            new.append(s)
        return updates, tuple(new)
    return GradientTransformation(init, update)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    return chain(scale_by_adam(b1, b2, eps, eps_root), scale_by_learning_rate(learning_rate))


def apply_updates(params, updates):
    return _map(lambda p, u: (torch.as_tensor(p) + u).to(torch.as_tensor(p).dtype), params, updates)


def linear_schedule(init_value, end_value, transition_steps, transition_begin=0):
    def schedule(count):
        frac = min(max(count - transition_begin, 0), transition_steps) / transition_steps if transition_steps > 0 else 1.0
        return init_value + (end_value - init_value) * frac
    return schedule


def cosine_decay_schedule(init_value, decay_steps, alpha=0.0, exponent=1.0):
    def schedule(count):
        c = min(count, decay_steps)
        cosine = 0.5 * (1 + math.cos(math.pi * c / decay_steps))
        return init_value * ((1 - alpha) * cosine ** exponent + alpha)
    return schedule


def join_schedules(schedules, boundaries):
    def schedule(count):
        i = sum(1 for b in boundaries if count >= b)
        return schedules[i](count - (boundaries[i - 1] if i > 0 else 0))
    return schedule


def warmup_cosine_decay_schedule(init_value, peak_value, warmup_steps, decay_steps, end_value=0.0, exponent=1.0):
    alpha = 0.0 if peak_value == 0.0 else end_value / peak_value
    return join_schedules([linear_schedule(init_value, peak_value, warmup_steps),
                           cosine_decay_schedule(peak_value, decay_steps - warmup_steps, alpha, exponent)], [warmup_steps])
