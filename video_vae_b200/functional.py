"""Autograd shell: one torch.autograd.Function per fused block of the hot path, each with a hand-written backward
made of libvvae kernels (video_vae_b200/ops.py).  PyTorch contributes tensors, the autograd tape between blocks and
a handful of scalar ops on [b,t]-sized bookkeeping tensors; all arithmetic on activations and parameters is ours.

Conventions
  * Parameters are fp32 (Flax layouts: Linear (in,out), conv (kt,kh,kw,Cin,Cout)).  When the compute dtype is bf16 a
    bf16 shadow copy is made once per parameter version (``shadow``) and used by forward and backward.
  * Parameter gradients are ACCUMULATED by the kernels straight into ``p.grad`` (fp32; created zero-filled on first
    use, or a view of a flat buffer when ``flatten_parameters`` was called) and the Functions return None for them,
    so autograd never runs an add kernel or allocates per-parameter gradient tensors.
  * Residual adds are fused: block Functions compute ``x + f(LN(x))`` and their backward emits a single dx.
"""
import math
import weakref

import torch
from torch.autograd import Function

from . import ops
from ._ffi import EPI_DSILU, EPI_RESIDUAL, EPI_SILU, require_device

# ------------------------------------------------------------------ parameter shadows / gradient buffers
_shadow_cache = {}  # id(param) -> (weakref(param), version, data_ptr, shadow tensor)
_flat_shadow_views = {}  # id(param) -> (weakref(param), bf16 view into FlatParams.shadow) (ddp.FlatParams.enable_bf16_shadow)
_grad_hooks = []  # callables(list_of_params) invoked after a backward Function has finished writing their grads
_decoder_done_hooks = []  # callables() invoked when backward reaches the latent: every decoder gradient is written


def shadow(p, dtype):
    """Compute-dtype copy of an fp32 parameter, refreshed when the parameter changes (in-place version bump)."""
    if p is None:
        return None
    if p.dtype == dtype:
        return p.detach()
    key = id(p)
    if dtype == torch.bfloat16:
        v = _flat_shadow_views.get(key)
        if v is not None and v[0]() is p:       # ids are recycled: the entry must belong to THIS parameter
            return v[1]
    ent = _shadow_cache.get(key)
    if (ent is not None and ent[0]() is p and ent[1] == p._version and ent[2] == p.data_ptr()
            and ent[3].dtype == dtype):
        return ent[3]
    t = ops.cast(p.detach(), dtype)
    _shadow_cache[key] = (weakref.ref(p, lambda _r, k=key: _shadow_cache.pop(k, None)), p._version, p.data_ptr(), t)
    return t


_param_epoch = [0]  # bumped whenever parameters change behind autograd's back (fused optimizer kernels)
_derived_cache = {}  # (id(param), key) -> (weakref(param), epoch, version, data_ptr, value)


def invalidate_shadows():
    _shadow_cache.clear()
    _param_epoch[0] += 1


def params_changed():
    """Called by the fused optimizer after it rewrote the flat parameter buffer in place.  The kernels write through raw
    pointers, so ``p._version`` does not move: per-parameter shadow copies made by ``shadow()`` (parameters outside any
    flat bf16 shadow) must be dropped as well, or they would keep serving the pre-update weights."""
    _shadow_cache.clear()
    _param_epoch[0] += 1


def derived(p, key, make):
    """Cache of a tensor derived from parameter ``p`` (e.g. a re-laid-out weight image); rebuilt when p changes."""
    k = (id(p), key)
    ent = _derived_cache.get(k)
    if (ent is not None and ent[0]() is p and ent[1] == _param_epoch[0] and ent[2] == p._version
            and ent[3] == p.data_ptr()):
        return ent[4]
    val = make()
    _derived_cache[k] = (weakref.ref(p, lambda _r, kk=k: _derived_cache.pop(kk, None)), _param_epoch[0], p._version,
                         p.data_ptr(), val)
    return val


def grad_buf(p):
    """fp32 accumulation buffer for d(loss)/dp; kernels add into it."""
    if p.grad is None:
        p.grad = ops.zeros_f32(p.shape, p.device)
    return p.grad


def _notify(params):
    if _grad_hooks:
        join_wgrad_lane()            # the hooks (gradient all-reduce) read these gradients on the current stream
    for h in _grad_hooks:
        h(params)


# ------------------------------------------------------------------ weight-gradient lane
# The weight-gradient kernels of a block (GEMMs with K = all tokens, conv3d wgrad) are consumed by nobody before the
# optimizer step, while the data-gradient kernels form the critical chain of backward.  They are launched on a second
# stream: every persistent kernel here ends with a partial wave (N = 768 dgrads: 5.2 waves of tiles on 74 CTA pairs) and
# a drain, during which the other stream's CTAs take the idle SMs.  Ordering: the lane waits for the current stream at
# the fork (its inputs are complete); the current stream waits for the lane two forks later (by then long finished, the
# wait is free) -- until then the lane's input tensors are kept alive here, so neither allocator reuse nor an in-place
# consumer can touch them -- and at the end of the backward pass (autograd callback), before anyone reads a gradient.
# Measured (bench.py --wgrad-lane, same box, 2 runs each): 70.5 / 71.0 ms per step with the lane against 70.0 / 70.3 ms
# without: two persistent kernels time-slicing the SMs buy nothing here (the step is power-capped, the kernels' CTAs
# cannot co-reside -- each takes ~200 KB of shared memory -- and the interleaving costs L2 locality).  OFF by default.
WGRAD_LANE = False                   # module switch (bench.py --wgrad-lane)
_lane = {"stream": None, "pending": [], "queued": False}


def join_wgrad_lane():
    """Current stream waits for everything launched on the weight-gradient lane so far."""
    if _lane["pending"]:
        main = torch.cuda.current_stream()
        for ev, _keep in _lane["pending"]:
            main.wait_event(ev)
        _lane["pending"].clear()
    _lane["queued"] = False


def wgrad_async(fn, *keep):
    """Run ``fn()`` (weight-gradient kernels reading ``keep``) on the lane, after all work already on the current stream."""
    if not WGRAD_LANE or not torch.cuda.is_available():
        fn()
        return
    if not _lane["queued"]:
        try:                           # join at the end of this backward pass (inside a graph capture: captured)
            torch.autograd.Variable._execution_engine.queue_callback(join_wgrad_lane)
            _lane["queued"] = True
        except RuntimeError:           # not inside a backward pass: no one to join for us
            fn()
            return
    main = torch.cuda.current_stream()
    if _lane["stream"] is None or _lane["stream"].device != main.device:
        _lane["stream"] = torch.cuda.Stream(device=main.device)
    side = _lane["stream"]
    pend = _lane["pending"]
    while len(pend) >= 2:
        ev, _keep = pend.pop(0)
        main.wait_event(ev)
    fork = torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        fn()
        done = torch.cuda.Event()
        done.record(side)
    pend.append((done, keep))


def _as_2d(x, D):
    x2 = x.reshape(-1, D)
    return x2 if x2.is_contiguous() else x2.contiguous()


def _rows_2d(x, D):
    """[.., D] -> [rows, D] keeping a row pitch > D when the leading dims are dense over it (no copy); else packed."""
    if x.stride(-1) == 1 and x.dim() >= 2:
        ld, ok = x.stride(-2), True
        for i in range(x.dim() - 2):
            ok = ok and x.stride(i) == x.stride(i + 1) * x.shape[i + 1]
        if ok and ld >= D:
            return torch.as_strided(x, (x.numel() // D, D), (ld, 1), x.storage_offset())
    return _as_2d(x, D)


def unet_input_pitch(C, dtype):
    """Elements per voxel of the U-Net's input map: ceil16(C) in bf16 (tensor-core conv channel blocks), else C."""
    if dtype == torch.bfloat16 and C % 16 != 0 and C % 4 == 0 and C < 32:
        return (C + 15) // 16 * 16
    return C


# ------------------------------------------------------------------ Linear
class LinearFn(Function):
    """y = x @ K + b (nnx.Linear); x may arrive in another float dtype (cast to the compute dtype first)."""

    @staticmethod
    def forward(ctx, x, kernel, bias, dtype):
        require_device()
        K, N = kernel.shape
        x2 = _as_2d(x, K)
        ctx.in_dtype = x2.dtype
        if x2.dtype != dtype:
            x2 = ops.cast(x2, dtype)
        w = shadow(kernel, dtype)
        y = ops.gemm(x2, w, bias=bias.detach() if bias is not None else None)
        ctx.save_for_backward(x2, kernel, bias)
        ctx.dtype = dtype
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, kernel, bias = ctx.saved_tensors
        K, N = kernel.shape
        dy2 = _as_2d(dy, N)
        if dy2.dtype != ctx.dtype:
            dy2 = ops.cast(dy2, ctx.dtype)
        ops.gemm(x2, dy2, transA=True, out=grad_buf(kernel), accumulate=True,
                 bsum=grad_buf(bias) if bias is not None else None)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dy2, shadow(kernel, ctx.dtype), transB=True)
            if dx.dtype != ctx.in_dtype:
                dx = ops.cast(dx, ctx.in_dtype)
            dx = dx.view(ctx.x_shape)
        _notify([kernel, bias])
        return dx, None, None, None


# ------------------------------------------------------------------ LayerNorm (standalone)
class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, scale, bias, dtype):
        require_device()
        D = x.shape[-1]
        x2 = _as_2d(x, D)
        if x2.dtype != dtype:
            x2 = ops.cast(x2, dtype)
        y, mean, rstd = ops.layernorm_fwd(x2, scale.detach() if scale is not None else None,
                                          bias.detach() if bias is not None else None)
        ctx.save_for_backward(x2, mean, rstd, scale, bias)
        ctx.x_shape, ctx.in_dtype = x.shape, x.dtype
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, scale, bias = ctx.saved_tensors
        dy2 = _as_2d(dy, x2.shape[1])
        dx = ops.layernorm_bwd(dy2, x2, mean, rstd, scale.detach() if scale is not None else None, None,
                               grad_buf(scale) if scale is not None else None,
                               grad_buf(bias) if bias is not None else None)
        if dx.dtype != ctx.in_dtype:
            dx = ops.cast(dx, ctx.in_dtype)
        _notify([scale, bias])
        return dx.view(ctx.x_shape), None, None, None


# ------------------------------------------------------------------ attention block
class AttnCfg:
    """Static description of one attention call: sequence geometry, heads, RoPE position rule, mask."""

    def __init__(self, geom, heads, hd, pos_div, pos_mod, mask, residual, dtype):
        self.geom, self.heads, self.hd = geom, heads, hd
        self.pos_div, self.pos_mod = pos_div, pos_mod
        self.mask, self.residual, self.dtype = mask, residual, dtype
        self.scale = 1.0 / math.sqrt(hd)


class AttnBlockFn(Function):
    """[x +] out_proj(attention(rope(qknorm(qkv_proj(LN(x)))))) -- train/layers.py:158-171 (+ residual of :214,220)."""

    @staticmethod
    def forward(ctx, x, cfg, ln_g, ln_b, w_qkv, b_qkv, q_scale, k_scale, w_o, b_o, cos, sin):
        require_device()
        D = x.shape[-1]
        dtp = cfg.dtype
        x2 = _as_2d(x, D)
        if x2.dtype != dtp:
            x2 = ops.cast(x2, dtp)
        Q = cfg.heads * cfg.hd
        h, mean, rstd = ops.layernorm_fwd(x2, ln_g.detach(), ln_b.detach())
        # projection + per-head QK-LayerNorm + RoPE in one call (tcgen05 path: in the GEMM epilogue)
        qkv, qk = ops.qkv_projection(h, shadow(w_qkv, dtp), b_qkv.detach(), q_scale.detach(), k_scale.detach(), cos, sin,
                                     cfg.heads, cfg.hd, cfg.pos_div, cfg.pos_mod)
        o, lse = ops.attn_fwd(cfg.geom, cfg.heads, cfg.hd, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], cfg.mask, cfg.scale)
        if cfg.residual:
            y = ops.gemm(o, shadow(w_o, dtp), bias=b_o.detach(), epilogue=EPI_RESIDUAL, aux_in=x2)
        else:
            y = ops.gemm(o, shadow(w_o, dtp), bias=b_o.detach())
        ctx.save_for_backward(x2, mean, rstd, h, qkv, qk, o, lse, ln_g, ln_b, w_qkv, b_qkv, q_scale, k_scale, w_o, b_o,
                              cos, sin)
        ctx.cfg = cfg
        ctx.x_shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (x2, mean, rstd, h, qkv, qk, o, lse, ln_g, ln_b, w_qkv, b_qkv, q_scale, k_scale, w_o, b_o, cos,
         sin) = ctx.saved_tensors
        cfg = ctx.cfg
        dtp = cfg.dtype
        D = x2.shape[1]
        Q = cfg.heads * cfg.hd
        dy2 = _as_2d(dy, D)
        # out projection; bias gradients ride along with the weight-gradient GEMMs (column sums of their dY tiles)
        gw_o, gb_o = grad_buf(w_o), grad_buf(b_o)
        wgrad_async(lambda: ops.gemm(o, dy2, transA=True, out=gw_o, accumulate=True, bsum=gb_o), o, dy2)
        d_o = ops.gemm(dy2, shadow(w_o, dtp), transB=True)
        # attention core -> dq | dk | dv written side by side
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(cfg.geom, cfg.heads, cfg.hd, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], o, lse, d_o, dqkv[:, :Q],
                     dqkv[:, Q:2 * Q], dqkv[:, 2 * Q:], cfg.mask, cfg.scale)
        ops.qknorm_rope_bwd_(dqkv, qkv, q_scale.detach(), k_scale.detach(), cos, sin, grad_buf(q_scale),
                             grad_buf(k_scale), cfg.heads, cfg.hd, cfg.pos_div, cfg.pos_mod)
        # qkv projection
        gw_qkv, gb_qkv = grad_buf(w_qkv), grad_buf(b_qkv)
        wgrad_async(lambda: ops.gemm(h, dqkv, transA=True, out=gw_qkv, accumulate=True, bsum=gb_qkv), h, dqkv)
        dh = ops.gemm(dqkv, shadow(w_qkv, dtp), transB=True)
        dx = ops.layernorm_bwd(dh, x2, mean, rstd, ln_g.detach(), dy2 if cfg.residual else None, grad_buf(ln_g),
                               grad_buf(ln_b), out=dh)
        _notify([ln_g, ln_b, w_qkv, b_qkv, q_scale, k_scale, w_o, b_o])
        return (dx.view(ctx.x_shape),) + (None,) * 11


# ------------------------------------------------------------------ MLP block
class MlpBlockFn(Function):
    """[x +] linear2(silu(linear1(LN(x)))) -- train/layers.py:191-196 (+ residual of :215,221)."""

    @staticmethod
    def forward(ctx, x, residual, dtype, ln_g, ln_b, w1, b1, w2, b2):
        require_device()
        D = x.shape[-1]
        x2 = _as_2d(x, D)
        if x2.dtype != dtype:
            x2 = ops.cast(x2, dtype)
        h, mean, rstd = ops.layernorm_fwd(x2, ln_g.detach(), ln_b.detach())
        u = torch.empty((x2.shape[0], w1.shape[1]), dtype=dtype, device=x2.device)
        a = ops.gemm(h, shadow(w1, dtype), bias=b1.detach(), epilogue=EPI_SILU, aux_out=u)
        if residual:
            y = ops.gemm(a, shadow(w2, dtype), bias=b2.detach(), epilogue=EPI_RESIDUAL, aux_in=x2)
        else:
            y = ops.gemm(a, shadow(w2, dtype), bias=b2.detach())
        ctx.save_for_backward(x2, mean, rstd, h, u, a, ln_g, ln_b, w1, b1, w2, b2)
        ctx.residual, ctx.dtype, ctx.x_shape = residual, dtype, x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, h, u, a, ln_g, ln_b, w1, b1, w2, b2 = ctx.saved_tensors
        dtp = ctx.dtype
        dy2 = _as_2d(dy, x2.shape[1])
        gw2, gb2, gw1, gb1 = grad_buf(w2), grad_buf(b2), grad_buf(w1), grad_buf(b1)
        wgrad_async(lambda: ops.gemm(a, dy2, transA=True, out=gw2, accumulate=True, bsum=gb2), a, dy2)
        du = ops.gemm(dy2, shadow(w2, dtp), transB=True, epilogue=EPI_DSILU, aux_in=u)
        wgrad_async(lambda: ops.gemm(h, du, transA=True, out=gw1, accumulate=True, bsum=gb1), h, du)
        dh = ops.gemm(du, shadow(w1, dtp), transB=True)
        dx = ops.layernorm_bwd(dh, x2, mean, rstd, ln_g.detach(), dy2 if ctx.residual else None, grad_buf(ln_g),
                               grad_buf(ln_b), out=dh)
        _notify([ln_g, ln_b, w1, b1, w2, b2])
        return (dx.view(ctx.x_shape),) + (None,) * 8


# ------------------------------------------------------------------ per-layer recompute (the reference's @nnx.remat)
class RecomputeFn(Function):
    """y = run(x) keeping ONLY x for backward; backward re-runs ``run`` (the forward kernels) and differentiates it.
    train/layers.py:209 and train/unet.py:44,76 wrap FactoredAttention / the U-Net blocks in ``@nnx.remat``; here it is a
    per-layer switch (default off: 180 GB of HBM hold the activations of the BASELINE shapes).
    ``params`` are passed only so that the output requires grad when x does not; their gradients are accumulated
    by the kernels into ``p.grad`` during the inner backward, as everywhere else."""

    @staticmethod
    def forward(ctx, x, run, *params):
        require_device()
        with torch.no_grad():
            y = run(x)
        ctx.run = run
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        with torch.enable_grad():
            xd = x.detach().requires_grad_(True)
            y = ctx.run(xd)
        torch.autograd.backward(y, dy)
        return (xd.grad, None) + (None,) * (len(ctx.needs_input_grad) - 2)


# ------------------------------------------------------------------ patch embedding / un-embedding
class PatchEmbedFn(Function):
    """rearrange -> cast -> LayerNorm -> Linear (train/layers.py:20-27).  The video needs no gradient."""

    @staticmethod
    def forward(ctx, video, P, dtype, ln_g, ln_b, kernel, bias):
        require_device()
        tok = ops.patchify(video, P, dtype)
        b, t, hw, D = tok.shape
        tok2 = tok.view(-1, D)
        h, mean, rstd = ops.layernorm_fwd(tok2, ln_g.detach(), ln_b.detach())
        y = ops.gemm(h, shadow(kernel, dtype), bias=bias.detach())
        ctx.save_for_backward(tok2, mean, rstd, h, ln_g, ln_b, kernel, bias)
        ctx.dtype = dtype
        return y.view(b, t, hw, kernel.shape[1])

    @staticmethod
    def backward(ctx, dy):
        tok2, mean, rstd, h, ln_g, ln_b, kernel, bias = ctx.saved_tensors
        dy2 = _as_2d(dy, kernel.shape[1])
        ops.gemm(h, dy2, transA=True, out=grad_buf(kernel), accumulate=True, bsum=grad_buf(bias))
        dh = ops.gemm(dy2, shadow(kernel, ctx.dtype), transB=True)
        ops.layernorm_bwd(dh, tok2, mean, rstd, ln_g.detach(), None, grad_buf(ln_g), grad_buf(ln_b), out=dh)
        _notify([ln_g, ln_b, kernel, bias])
        return (None,) * 7


class UnembedFn(Function):
    """Linear -> Linear(x u) -> pixel shuffle -> per-pixel Linear(C*u -> C)  (train/layers.py:45-55).
    Returns (features [b,t,H,W,C*u], rgb [b,t,H,W,C])."""

    @staticmethod
    def forward(ctx, x, geo, dtype, wl, bl, wu, bu, wd, bd):
        require_device()
        b, t, H, W, P, CU = geo
        D = x.shape[-1]
        x2 = _as_2d(x, D)
        y1 = ops.gemm(x2, shadow(wl, dtype), bias=bl.detach())
        y2 = ops.gemm(y1, shadow(wu, dtype), bias=bu.detach())
        # bf16: the features are the U-Net's input, whose tensor-core convolutions gather 16-channel blocks: produce them
        # at that pitch (zeroed pad channels) instead of re-packing 268 MB later; the returned tensor is the [.., :CU] view
        ld = unet_input_pitch(CU, dtype)
        feats = ops.pixel_shuffle(y2, b * t, H, W, CU, P, to_tokens=False, vox_ld=ld)
        f2 = feats.view(-1, ld)[:, :CU]
        rgb = ops.gemm(f2, shadow(wd, dtype), bias=bd.detach())
        ctx.save_for_backward(x2, y1, feats, wl, bl, wu, bu, wd, bd)
        ctx.geo, ctx.dtype, ctx.x_shape = geo, dtype, x.shape
        Cc = wd.shape[1]
        return feats.view(b, t, H, W, ld)[..., :CU], rgb.view(b, t, H, W, Cc)

    @staticmethod
    def backward(ctx, dfeats, drgb):
        x2, y1, feats, wl, bl, wu, bu, wd, bd = ctx.saved_tensors
        b, t, H, W, P, CU = ctx.geo
        dtp = ctx.dtype
        Cc = wd.shape[1]
        ld = feats.shape[-1]
        f2 = feats.view(-1, ld)[:, :CU]
        if dfeats is not None:                         # [.., :CU] view of a pitched buffer (U-Net backward) or packed
            dfeats = _rows_2d(dfeats, CU)
        if drgb is not None:
            drgb2 = _as_2d(drgb, Cc)
            ops.gemm(f2, drgb2, transA=True, out=grad_buf(wd), accumulate=True)
            ops.colsum_accum(drgb2, grad_buf(bd))
            if dfeats is not None:
                df = ops.gemm(drgb2, shadow(wd, dtp), transB=True, epilogue=EPI_RESIDUAL, aux_in=dfeats)
            else:
                df = ops.gemm(drgb2, shadow(wd, dtp), transB=True)
            dy2 = ops.pixel_shuffle(df, b * t, H, W, CU, P, to_tokens=True)
        elif dfeats.stride(0) != CU and dfeats.stride(0) == unet_input_pitch(CU, dtp):
            dy2 = ops.pixel_shuffle(dfeats, b * t, H, W, CU, P, to_tokens=True, vox_ld=dfeats.stride(0))
        else:
            dy2 = ops.pixel_shuffle(_as_2d(dfeats, CU), b * t, H, W, CU, P, to_tokens=True)
        ops.gemm(y1, dy2, transA=True, out=grad_buf(wu), accumulate=True, bsum=grad_buf(bu))
        dy1 = ops.gemm(dy2, shadow(wu, dtp), transB=True)
        ops.gemm(x2, dy1, transA=True, out=grad_buf(wl), accumulate=True, bsum=grad_buf(bl))
        dx = ops.gemm(dy1, shadow(wl, dtp), transB=True)
        _notify([wl, bl, wu, bu, wd, bd])
        return (dx.view(ctx.x_shape),) + (None,) * 8


# ------------------------------------------------------------------ encoder head, reparameterisation, loss
class EncoderHeadFn(Function):
    """mean / log-variance heads and the Gumbel-sigmoid frame gate (train/model.py:53-59, train/layers.py:238-252).
    Returns (mean [b,t,hw,Dl], logvar, selection [b,t,1,1] fp32).  ``train`` = False / True, or the string "prob" for
    the rl_model variant (train/rl_model.py:59): selection = sigmoid(logit) itself, no Gumbel noise, no rounding."""

    @staticmethod
    def forward(ctx, x, dtype, train, temperature, u, seed, offset, wm, bm, wv, bv, w1, b1, w2, b2):
        require_device()
        b, t, hw, D = x.shape
        x2 = _as_2d(x, D)
        mean = ops.gemm(x2, shadow(wm, dtype), bias=bm.detach())
        a = ops.gemm(x2, shadow(wv, dtype), bias=bv.detach())
        lv = ops.softplus_log_fwd(a)
        s1 = ops.gemm(mean, shadow(w1, dtype), bias=b1.detach())                  # [N,1]
        if w2.shape[0] != hw:
            raise ValueError(f"selection_layer2 expects {w2.shape[0]} spatial tokens, got {hw}")
        prob_mode = train == "prob"
        logit, p, sel = ops.selection_fwd(s1, w2.detach().reshape(-1), b2.detach(), u, seed, offset,
                                          False if prob_mode else train, temperature, b * t, hw)
        if prob_mode:
            sel = p
        ctx.save_for_backward(x2, mean, a, s1, p, wm, bm, wv, bv, w1, b1, w2, b2)
        ctx.dtype, ctx.train, ctx.temperature, ctx.shape = dtype, train, temperature, (b, t, hw, D)
        Dl = wm.shape[1]
        return mean.view(b, t, hw, Dl), lv.view(b, t, hw, Dl), sel.view(b, t, 1, 1)

    @staticmethod
    def backward(ctx, dmean, dlv, dsel):
        x2, mean, a, s1, p, wm, bm, wv, bv, w1, b1, w2, b2 = ctx.saved_tensors
        b, t, hw, D = ctx.shape
        dtp = ctx.dtype
        Dl = wm.shape[1]
        dmean2 = _as_2d(dmean, Dl) if dmean is not None else None
        if dmean2 is not None and dmean2.dtype != dtp:
            dmean2 = ops.cast(dmean2, dtp)
        if dsel is not None:
            # round_ste passes the gradient through; d sigmoid((a+g)/T) = p(1-p)/T          [b*t scalars: torch glue]
            da_sel = dsel.reshape(-1).float() * p * (1.0 - p) / ctx.temperature           # [bt]
            s1f = s1.view(b * t, hw)
            w2f = w2.detach().reshape(-1)
            ds1 = (da_sel[:, None] * w2f[None, :]).to(dtp).reshape(-1, 1).contiguous()    # [N,1]
            grad_buf(w2).view(-1).add_((s1f.float() * da_sel[:, None]).sum(0))
            grad_buf(b2).view(-1).add_(da_sel.sum())
            ops.gemm(mean, ds1, transA=True, out=grad_buf(w1), accumulate=True)
            ops.colsum_accum(ds1, grad_buf(b1))
            if dmean2 is not None:
                dmean2 = ops.gemm(ds1, shadow(w1, dtp), transB=True, epilogue=EPI_RESIDUAL, aux_in=dmean2)
            else:
                dmean2 = ops.gemm(ds1, shadow(w1, dtp), transB=True)
        dx = None
        if dmean2 is not None:
            ops.gemm(x2, dmean2, transA=True, out=grad_buf(wm), accumulate=True, bsum=grad_buf(bm))
            dx = ops.gemm(dmean2, shadow(wm, dtp), transB=True)
        if dlv is not None:
            dlv2 = _as_2d(dlv, Dl)
            if dlv2.dtype != dtp:
                dlv2 = ops.cast(dlv2, dtp)
            da = ops.softplus_log_bwd(dlv2, a)
            ops.gemm(x2, da, transA=True, out=grad_buf(wv), accumulate=True, bsum=grad_buf(bv))
            if dx is not None:
                dx = ops.gemm(da, shadow(wv, dtp), transB=True, epilogue=EPI_RESIDUAL, aux_in=dx)
            else:
                dx = ops.gemm(da, shadow(wv, dtp), transB=True)
        if dx is None:
            dx = torch.zeros_like(x2)
        _notify([wm, bm, wv, bv, w1, b1, w2, b2])
        return (dx.view(b, t, hw, D),) + (None,) * 14


class ReparamGateFn(Function):
    """z = mean + eps*exp(lv/2); c = fill*(1-sel) + z*sel  (train/model.py:124-133).
    Returns (c fp32 -- what the reference returns --, c in the compute dtype for the decoder)."""

    @staticmethod
    def forward(ctx, mean, logvar, sel, fill, eps, seed, offset, train):
        require_device()
        b, t, hw, Dl = mean.shape
        lowp = mean.dtype != torch.float32
        c32, cT, eps_used = ops.reparam_gate_fwd(mean.contiguous(), logvar.contiguous(), eps, seed, offset,
                                                 sel.reshape(-1).contiguous(), fill.detach().reshape(-1), hw, train, lowp)
        ctx.save_for_backward(mean, logvar, sel, fill, eps_used if eps_used is not None else mean.new_empty(0))
        ctx.train, ctx.hw = train, hw
        return c32, (cT if lowp else None)

    @staticmethod
    def backward(ctx, dc32, dcT):
        for h in _decoder_done_hooks:      # the decoder's backward (incl. its first Linear) has been enqueued
            h()
        mean, logvar, sel, fill, eps = ctx.saved_tensors
        if dcT is None and dc32 is None:
            return (None,) * 8
        if dcT is not None and dc32 is not None and dcT.data_ptr() != dc32.data_ptr():
            dc = dcT + dc32.to(dcT.dtype)
        else:
            dc = dcT if dcT is not None else dc32
        if dc.dtype != mean.dtype:
            dc = ops.cast(dc.contiguous(), mean.dtype)
        dsel = ops.zeros_f32((sel.numel(),), mean.device)
        dmean, dlogvar = ops.reparam_gate_bwd(dc.contiguous(), mean, logvar, eps if ctx.train else None,
                                              sel.reshape(-1).contiguous(), fill.detach().reshape(-1), None, None,
                                              grad_buf(fill).view(-1), dsel, ctx.hw, ctx.train)
        _notify([fill])
        return dmean, (dlogvar if ctx.train else None), dsel.view(sel.shape), None, None, None, None, None


def magnify_negatives(x, rate):
    return torch.where(x < 0, x * rate, x)


class VaeLossFn(Function):
    """loss_fn of train/legacy/training_loop_adversarial.py:90-124 (+ optional MAE term, rl_nonadversarial.py:114-117).
    Returns (loss, MSE, MAE, selection_loss, kl_loss, kept_frame_density) -- only ``loss`` is differentiable."""

    @staticmethod
    def forward(ctx, video, recon, sel, logvar, mean, mask_bt, hp):
        require_device()
        B, T = mask_bt.shape
        dev = recon.device
        m = mask_bt.to(torch.float32)
        seq = torch.clamp(m.sum(dim=1), min=1.0)
        inv_len = (1.0 / seq).contiguous()
        frame_w = (m * inv_len[:, None]).reshape(-1).contiguous()
        sums = ops.zeros_f32((3,), dev)
        video = video.contiguous()
        recon = recon.contiguous()
        ops.recon_loss_fwd(video, recon, m.reshape(-1).contiguous(), inv_len, sums[:2])
        hw = mean.shape[2]
        ops.kl_fwd(mean.contiguous(), logvar.contiguous(), frame_w, sums[2:], hw)
        count = float(B * (video.numel() // (B * T)))
        mse, mae = sums[0] / count, sums[1] / count
        kl = sums[2] / float(mean.numel())
        density = (sel.reshape(B, T).to(torch.float32) * m).sum(dim=1) * inv_len
        diff = density - 1.0 / hp["max_compression_rate"]
        mag = magnify_negatives(diff, hp["magnify_negatives_rate"])
        sel_loss = torch.square(mag).mean()
        g4 = hp.get("gamma4", 0.0)
        loss = mse + hp["gamma1"] * sel_loss + hp["gamma2"] * kl + g4 * mae
        ctx.save_for_backward(video, recon, sel, logvar, mean, m, inv_len, frame_w, diff)
        ctx.hp, ctx.count = hp, count
        dens = density.mean()
        ctx.mark_non_differentiable(mse, mae, sel_loss, kl, dens)
        return loss, mse, mae, sel_loss, kl, dens

    @staticmethod
    def backward(ctx, gl, *unused):
        video, recon, sel, logvar, mean, m, inv_len, frame_w, diff = ctx.saved_tensors
        hp = ctx.hp
        B, T = m.shape
        # the upstream gradient stays on the device (a kernel argument), so the step never waits for the host
        g = gl.detach().reshape(1).to(torch.float32)
        drecon = ops.recon_loss_bwd(video, recon, m.reshape(-1).contiguous(), inv_len, 1.0, hp.get("gamma4", 0.0),
                                    1.0 / ctx.count, gscale=g)
        dmean, dlogvar = ops.kl_bwd(mean, logvar, frame_w, hp["gamma2"] / float(mean.numel()), mean.shape[2], gscale=g)
        rate = hp["magnify_negatives_rate"]
        slope = torch.where(diff < 0, torch.full_like(diff, rate), torch.ones_like(diff))
        ddens = g * ((hp["gamma1"] / B) * 2.0) * (diff * slope) * slope                    # [B]
        dsel = (ddens[:, None] * m * inv_len[:, None]).reshape(sel.shape).to(sel.dtype)
        return None, drecon, dsel, dlogvar, dmean, None, None
