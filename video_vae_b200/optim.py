"""Host side of the optimizer row: the learning-rate schedule of train/rl_nonadversarial.py:241-247.

``optax.warmup_cosine_decay_schedule(init_value, peak_value, warmup_steps, decay_steps, end_value)`` as a plain Python
function of the update count (the first update has count 0, so the warm-up starts with a zero-size step, exactly as
optax's ``scale_by_learning_rate`` does).  The arithmetic on the parameters (global-norm clip + Adam) is the fused
kernel pair ``vvae_sumsq_f32`` / ``vvae_adam_step`` driven by ``ddp.FlatAdam``; pass the schedule as its ``lr``.
"""
import math


def warmup_cosine_decay_schedule(init_value, peak_value, warmup_steps, decay_steps, end_value=0.0, exponent=1.0):
    warmup_steps = float(warmup_steps)
    cos_steps = float(decay_steps) - warmup_steps
    alpha = 0.0 if peak_value == 0 else end_value / peak_value

    def schedule(count):
        count = float(count)
        if count < warmup_steps:
            return init_value + (peak_value - init_value) * (count / warmup_steps if warmup_steps > 0 else 1.0)
        c = min(count - warmup_steps, cos_steps)
        cosine = 0.5 * (1.0 + math.cos(math.pi * c / cos_steps)) if cos_steps > 0 else 0.0
        return peak_value * ((1.0 - alpha) * cosine ** exponent + alpha)

    return schedule


def reference_schedule(batch_size, learning_rate=2e-5, decay_steps=1_000_000):
    """The production schedule: LEARNING_RATE 2e-5, WARMUP_STEPS = 20000 // sqrt(batch), DECAY_STEPS 1e6, end lr / 10
    (train/rl_nonadversarial.py:46,51-52,241-247)."""
    return warmup_cosine_decay_schedule(0.0, learning_rate, 20000 // math.sqrt(batch_size), decay_steps, learning_rate / 10)
