"""Oracle restatement of the reference optimizer (CPU torch, TEST INFRASTRUCTURE ONLY).

The reference builds ``optax.chain(optax.clip_by_global_norm(1.0), optax.adam(learning_rate=schedule_fn))`` with
``schedule_fn = optax.warmup_cosine_decay_schedule(0.0, LEARNING_RATE, WARMUP_STEPS, DECAY_STEPS, LEARNING_RATE / 10)``
(train/rl_nonadversarial.py:241-253; constants :46,51-52: LEARNING_RATE 2e-5, WARMUP_STEPS 20000 // sqrt(batch),
DECAY_STEPS 1e6).  optax (pinned 0.2.6, claude_distributed/requirements.txt) is not vendored and cannot be installed
offline, so its published algorithms are restated here [optax-recall]:

* clip_by_global_norm(c): g <- g                      if ||g||_2 < c
                          g <- g / ||g||_2 * c        otherwise            (norm over ALL leaves)
* scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0): m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ;
  u = (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps),  t = 1, 2, ...
* scale_by_learning_rate(schedule): p <- p - schedule(count) * u with count = 0 for the FIRST update
  (the warm-up therefore starts with a step of size init_value = 0).
* warmup_cosine_decay_schedule = join_schedules([linear 0 -> peak over warmup_steps,
  cosine_decay(peak, decay_steps - warmup_steps, alpha = end / peak)], [warmup_steps]).

Parity unpinned against optax itself; pinned against torch.optim.Adam + clip_grad_norm_ (tests/test_oracle.py), against a
second restatement that keeps optax's own structure (oracle/jaxshim/optax.py, tests/test_jaxshim_cpu.py) and against the
reference's own train_step + optimizer construction run on that look-alike for six updates
(tests/golden/refshim_rltrain_small_float32.npz, tests/test_jax_golden.py).
"""
import math

import torch


def warmup_cosine_decay_schedule(init_value, peak_value, warmup_steps, decay_steps, end_value=0.0, exponent=1.0):
    warmup_steps = float(warmup_steps)
    cos_steps = float(decay_steps) - warmup_steps
    alpha = 0.0 if peak_value == 0 else end_value / peak_value

    def schedule(count):
        count = float(count)
        if count < warmup_steps:
            frac = count / warmup_steps if warmup_steps > 0 else 1.0
            return init_value + (peak_value - init_value) * frac
        c = min(count - warmup_steps, cos_steps)
        cosine = 0.5 * (1.0 + math.cos(math.pi * c / cos_steps)) if cos_steps > 0 else 0.0
        return peak_value * ((1.0 - alpha) * cosine ** exponent + alpha)

    return schedule


class ClipAdam:
    """optax.chain(clip_by_global_norm(clip), adam(lr)) over a list of fp32 tensors (updated in place)."""

    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8, clip=1.0):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps, self.clip = lr, b1, b2, eps, clip
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.count = 0

    def step(self, grads):
        gn = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
        scale = 1.0 if gn < self.clip else self.clip / gn
        lr = self.lr(self.count) if callable(self.lr) else self.lr
        self.count += 1
        t = self.count
        bc1, bc2 = 1.0 - self.b1 ** t, 1.0 - self.b2 ** t
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                g = g * scale
                m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                p.sub_(lr * (m / bc1) / (torch.sqrt(v / bc2) + self.eps))
        return gn
