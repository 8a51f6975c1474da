"""CPU oracle for the video-VAE hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch on the CPU, the algorithm of the
reference's encoder/decoder hot path (train/layers.py, train/model.py,
train/unet.py and the loss of train/legacy/training_loop_adversarial.py:90-124,
MAE term train/rl_nonadversarial.py:114-117).  The arithmetic of the reference
lives in un-vendored third-party packages (flax 0.12.4 nnx.Linear / LayerNorm /
GroupNorm / Conv / ConvTranspose / max_pool, jax 0.9.0.1
jax.nn.dot_product_attention / softplus / sigmoid -- pins in
claude_distributed/requirements.txt:13,16); their published defaults are
restated in oracle/nn.py, each function citing the reference call site.

PARITY UNPINNED: the reference ships no golden vectors or known-answer tests for
this path and JAX/Flax cannot be installed or run in this environment (no
network), so the oracle is checked only against (a) an independent float64 numpy
restatement of every primitive (oracle/np_ref.py) and (b) the reference's own
property tests (masked == truncated attention, batch isolation, shapes, binary
gate, STE gradient, loss decrease).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (video_vae_b200/) never
does, and has no CPU path at all.
"""
from .rng import Rngs  # noqa: F401
