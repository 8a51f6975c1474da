"""CUDA-graph capture of the training step.

One step of the production model is ~1200 kernel launches; enqueuing them from Python costs about as much wall time
as the B200 needs to execute them, so the whole forward + loss + backward is captured ONCE into a CUDA graph and
replayed with a single launch per step.  What varies between steps lives in static device buffers the graph reads:
the input clip and mask (copied in before the replay) and the Philox draws of the step (materialised by
``vvae_philox_fill`` from the same (seed, offset) pairs the fused kernels would use, so a replay is bit-identical to
the eager step).  The optimizer (and, for N > 1, the gradient all-reduce) runs outside the graph: the bias-corrected
Adam constants change every step and are passed by value.

Capture BEFORE running eager backward passes in the same process: after an eager ``loss.backward()`` the autograd
engine's stream bookkeeping for the leaves makes a later capture fail with "dependency created on uncaptured work in
another stream" (tests and scripts/run_configs.py capture first, then run their eager comparisons).  Nothing on the
captured path may read device data from the host: ``repeat_interleave`` with an integer count and the backward of
``prod`` both do, which is why rl_model / rl_losses avoid them.
"""
import torch

from . import functional as F_
from . import ops
from .losses import DEFAULT_HPARAMS, loss_fn


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
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            F_.invalidate_shadows()                      # every derived weight image is rebuilt INSIDE the graph
            hook = None
            if self.decoder_done is not None:
                hook = lambda: self.decoder_done.record(torch.cuda.current_stream())   # noqa: E731
                F_._decoder_done_hooks.append(hook)
            try:
                with torch.cuda.graph(self.graph):
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
