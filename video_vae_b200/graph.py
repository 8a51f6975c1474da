"""CUDA-graph capture of the training step.

One step of the production model is ~1200 kernel launches; enqueuing them from Python costs about as much wall time
as the B200 needs to execute them, so the whole forward + loss + backward is captured ONCE into a CUDA graph and
replayed with a single launch per step.  What varies between steps lives in static device buffers the graph reads:
the input clip and mask (copied in before the replay) and the Philox draws of the step (materialised by
``vvae_philox_fill`` from the same (seed, offset) pairs the fused kernels would use, so a replay is bit-identical to
the eager step).  The optimizer (and, for N > 1, the gradient all-reduce) runs outside the graph: the bias-corrected
Adam constants change every step and are passed by value.

Capture works at any point of a process, also after eager ``loss.backward()`` calls.  (The trap, found in round 1 and
diagnosed in round 2: every leaf's AccumulateGrad node is bound to the stream that was current when it was created and
lives as long as any old autograd graph references it; the captured backward then synchronises the capture stream with
that uncaptured stream -> "dependency created on uncaptured work in another stream".  The capture therefore runs on
PRIVATE LEAVES: for the duration of warm-up + capture every parameter of the model is replaced by a fresh
``nn.Parameter`` over the same storage and the same gradient view, whose AccumulateGrad nodes are born on the capture
stream; the recorded kernels only know addresses, so replays serve the original parameters.)  Nothing on the
captured path may read device data from the host: ``repeat_interleave`` with an integer count and the backward of
``prod`` both do, which is why rl_model / rl_losses avoid them.
"""
import contextlib
import weakref

import torch
from torch import nn

from . import functional as F_
from . import ops
from .losses import DEFAULT_HPARAMS, loss_fn


@contextlib.contextmanager
def _private_leaves(model, reducer=None):
    """Swap every parameter of ``model`` for a fresh leaf over the same storage / gradient buffer (see module docstring)."""
    swapped = []
    for mod in model.modules():
        for name, p in list(mod._parameters.items()):
            if p is None:
                continue
            twin = nn.Parameter(p.data, requires_grad=p.requires_grad)
            twin.grad = p.grad
            ent = F_._flat_shadow_views.get(id(p))
            if ent is not None and ent[0]() is p:
                F_._flat_shadow_views[id(twin)] = (weakref.ref(twin), ent[1])
            if reducer is not None and id(p) in reducer.bucket_of:
                reducer.bucket_of[id(twin)] = reducer.bucket_of[id(p)]
            mod._parameters[name] = twin
            swapped.append((mod, name, p, twin))
    try:
        yield
    finally:
        for mod, name, p, twin in swapped:
            mod._parameters[name] = p
            if p.grad is None and twin.grad is not None:      # a gradient buffer first created during the capture
                p.grad = twin.grad
            F_._flat_shadow_views.pop(id(twin), None)
            if reducer is not None:
                reducer.bucket_of.pop(id(twin), None)


class GraphedTrainStep:
    """fwd + loss + bwd of ``model`` on a fixed (batch, frames, size) shape as one CUDA graph."""

    def __init__(self, model, flat, video_like, mask_like, hparams=None, warmup=2, reducer=None, mark_decoder_done=False):
        """reducer: an optional ddp.GradAllReducer.  Its bucketed NCCL all-reduces are then captured INSIDE the graph on
        the reducer's side stream (fork after a bucket's last gradient kernel, join before the graph ends), so the
        gradient exchange overlaps the rest of backward exactly as in eager mode."""
        self.model, self.flat, self.reducer = model, flat, reducer
        self.hp = dict(DEFAULT_HPARAMS if hparams is None else hparams)
        dev = video_like.device
        b, t = mask_like.shape
        hw = (video_like.shape[2] // model.encoder.patch_embedding.patch_size) * \
             (video_like.shape[3] // model.encoder.patch_embedding.patch_size)
        lat = model.fill_token.shape[-1]
        self.video = torch.empty_like(video_like)
        self.mask = torch.empty_like(mask_like)
        self.noise = torch.empty(b, t, hw, lat, dtype=torch.float32, device=dev)
        self.gate_u = torch.empty(self._gate_shape(b, t), dtype=torch.float32, device=dev)
        self.video.copy_(video_like)
        self.mask.copy_(mask_like)
        ops.philox_fill_(self.noise, 0, 1 << 40, "normal")
        ops.philox_fill_(self.gate_u, 0, 2 << 40, "uniform")
        self.graph = torch.cuda.CUDAGraph()
        # external event recorded INSIDE the graph when backward reaches the latent (all decoder gradients written): a
        # communication stream can start reducing the decoder's gradients while the encoder's backward still runs
        self.decoder_done = torch.cuda.Event(external=True) if mark_decoder_done else None
        self.loss = None
        self.aux = None
        prof, ops.PROFILE = ops.PROFILE, None            # event timing cannot be captured
        try:
            with _private_leaves(model, reducer):
                side = torch.cuda.Stream()               # warm-up AND capture on this one stream
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(warmup):
                        self._body()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                F_.invalidate_shadows()                  # every derived weight image is rebuilt INSIDE the graph
                hook = None
                if self.decoder_done is not None:
                    hook = lambda: self.decoder_done.record(torch.cuda.current_stream())   # noqa: E731
                    F_._decoder_done_hooks.append(hook)
                try:
                    with torch.cuda.graph(self.graph, stream=side):
                        self.loss, self.aux = self._body()
                finally:
                    if hook is not None:
                        F_._decoder_done_hooks.remove(hook)
        finally:
            ops.PROFILE = prof

    # -- what differs between the model variants: the gate's uniform draws and the loss
    def _gate_shape(self, b, t):
        return (b, t, 1)                                # Gumbel-sigmoid gate: one draw per frame (train/model.py:57-59)

    def _loss(self):
        return loss_fn(self.model, self.video, self.mask[:, None, None, :], self.mask, None, self.hp, train=True,
                       noise=self.noise, gumbel_u=self.gate_u)

    def _refresh_draws(self, rngs):
        """Same draws, same order as the eager path: the encoder's gate first, then the reparameterisation."""
        seed, off = rngs.sampling()
        ops.philox_fill_(self.gate_u, seed, off, "uniform")
        seed, off = rngs.sampling()
        ops.philox_fill_(self.noise, seed, off, "normal")

    def _body(self):
        self.flat.zero_grad()
        if self.reducer is not None:
            self.reducer.start_step()
        loss, aux = self._loss()
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish_step()
        return loss.detach(), {k: v.detach() for k, v in aux.items()}

    def __call__(self, video, mask, rngs):
        self.video.copy_(video, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        self._refresh_draws(rngs)
        self.graph.replay()
        return self.loss


class GraphedRLTrainStep(GraphedTrainStep):
    """The RL step (rl_model.VideoVAE + rl_losses.loss_fn, train/rl_nonadversarial.py:100-198) as one CUDA graph.
    ``perceptual_loss_fn`` / ``vgg_params`` as in the reference's loss_fn (None drops the term)."""

    def __init__(self, model, flat, video_like, mask_like, hparams=None, perceptual_loss_fn=None, vgg_params=None, **kw):
        from .rl_losses import DEFAULT_HPARAMS as RL_HP
        self.perceptual_loss_fn, self.vgg_params = perceptual_loss_fn, vgg_params
        super().__init__(model, flat, video_like, mask_like, dict(RL_HP if hparams is None else hparams), **kw)

    def _gate_shape(self, b, t):
        return (2 * b, t)                               # Bernoulli keep-mask: one draw per frame of the doubled batch

    def _loss(self):
        from .rl_losses import loss_fn as rl_loss_fn
        return rl_loss_fn(self.model, self.video, self.mask[:, None, None, :], self.mask, None, self.hp,
                          self.perceptual_loss_fn, self.vgg_params, train=True, noise=self.noise, bernoulli_u=self.gate_u)

    def _refresh_draws(self, rngs):
        """Eager order of rl_model.VideoVAE.forward: the Gaussian latent noise first, then the keep-mask draws."""
        seed, off = rngs.sampling()
        ops.philox_fill_(self.noise, seed, off, "normal")
        seed, off = rngs.sampling()
        ops.philox_fill_(self.gate_u, seed, off, "uniform")
