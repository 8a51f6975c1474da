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

PARITY STATUS: pinned against the reference's OWN CODE, unpinned against real
JAX.  The reference ships no golden vectors and JAX/Flax cannot be installed
here (no network).  What pins the oracle: (a) tests/golden/refshim_*.npz --
train/model.py, rl_model.py, layers.py, unet.py and the loss functions of
train/legacy/training_loop_adversarial.py and train/rl_nonadversarial.py
executed UNMODIFIED from /root/reference on oracle/jaxshim (thin jax / flax.nnx
look-alikes on CPU torch; generator tests/golden/make_golden_jax.py --shim,
consumers tests/test_jax_golden.py), which fixes every line the reference wrote;
(b) an independent float64 numpy restatement of every third-party primitive
(oracle/np_ref.py), against which both oracle/nn.py and the shim's primitives
are checked; (c) the reference's own checks (train/attention_mask_tests.py run
unmodified on the shim; masked == truncated, batch isolation, shapes, binary
gate, STE gradient, loss decrease).  What stays unpinned: flax's and jax's own
primitive semantics, restated (twice, independently) from their published
algorithms -- a fixture from real JAX (same generator without --shim, file
jax_*.npz, same consumers) closes that.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (video_vae_b200/) never
does, and has no CPU path at all.
"""
from .rng import Rngs  # noqa: F401
