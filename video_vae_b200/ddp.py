"""Data-parallel plumbing: flat fp32 parameter / gradient storage, bucketed gradient all-reduce overlapped with
backward, and the fused Adam step.

Mirrors the DP semantics of claude_distributed/distributed_train.py:107-109,189-196,378-380 (parameters replicated,
batch sharded, loss = mean over the global batch => gradient all-reduce-mean) with one process per GPU and NCCL over
NVLink (torch.distributed is the plumbing).  The only collective on the path is the gradient all-reduce; it is issued
per bucket, in the order backward finishes the buckets (decoder first), on a dedicated stream, so it hides under the
remaining backward kernels.  The 1/world_size scaling is folded into the optimizer's ``grad_scale``.
"""
import weakref

import torch
import torch.distributed as dist

from . import functional as F_
from . import ops


def shard_batch(x, rank=None, world=None):
    """Rows of a GLOBAL batch this rank owns: axis 0 split evenly over the ranks in rank order (the reference shards the
    batch over its 1-D ('data',) mesh, row block d on device d: claude_distributed/distributed_train.py:107-109,189-196;
    test_training_loop.py:221-233).  The global batch must divide by the world size, as the reference requires."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    n = x.shape[0]
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by the world size {world}")
    per = n // world
    return x[rank * per:(rank + 1) * per]


class NativeComm:
    """The C ABI's own communicator (``vvae_comm_*``, csrc/comm.cu: NCCL resolved with dlopen) for hosts without
    torch.distributed -- what the jax.ffi binding of INTEGRATION.md would drive.  One per process / GPU.

    ``exchange`` carries rank 0's 128-byte rendezvous token to every rank: a callable ``bytes | None -> bytes`` (rank 0
    passes the token, the others pass None and receive it).  With torch.distributed initialised the default exchange is
    ``broadcast_object_list`` (any backend: the token is host data); ``from_env`` transports it through a file.
    The package's own training path keeps using torch.distributed (``FlatParams.all_reduce_grads``); pass ``comm=`` there
    to route the same collective through this class instead."""

    TOKEN_BYTES = 128

    def __init__(self, rank, world, exchange=None):
        import ctypes as C
        from . import _ffi
        self._ffi, self._C = _ffi, C
        self.rank, self.world, self._h = int(rank), int(world), None
        _ffi.require_device()                                    # no CPU path: fail before any rank blocks in the rendezvous
        token = None
        if self.rank == 0:
            buf = (C.c_char * self.TOKEN_BYTES)()
            self._check(_ffi.lib.vvae_comm_unique_id(C.cast(buf, C.c_void_p)), "vvae_comm_unique_id")
            token = bytes(buf)
        if self.world > 1:
            token = (exchange or self._torch_exchange)(token)
        if not isinstance(token, (bytes, bytearray)) or len(token) != self.TOKEN_BYTES:
            raise ValueError("NativeComm: the exchange must return rank 0's 128-byte token on every rank")
        buf = (C.c_char * self.TOKEN_BYTES).from_buffer_copy(bytes(token))
        h = C.c_void_p()
        self._check(_ffi.lib.vvae_comm_init(C.byref(h), C.cast(buf, C.c_void_p), self.rank, self.world), "vvae_comm_init")
        self._h = h

    def _check(self, rc, what):          # not _ffi.check: collectives are not counted as kernel launches of this library
        if rc != 0:
            raise self._ffi.VvaeError(f"{what} failed (status {rc}): {self._ffi.lib.vvae_last_error().decode()}")

    @staticmethod
    def _torch_exchange(token):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NativeComm: pass exchange= (no torch.distributed process group to carry the token)")
        box = [token]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    @classmethod
    def file_exchange(cls, path, timeout_s=300.0):
        """A token exchange through the file ``path`` on a filesystem all ranks see: rank 0 writes it atomically, the
        others poll for it.  Use a fresh path per job (a stale file of an earlier job would be picked up)."""
        import os
        import time

        def exchange(token):
            if token is not None:
                with open(path + ".tmp", "wb") as f:
                    f.write(token)
                os.replace(path + ".tmp", path)
                return token
            t0 = time.monotonic()
            while not (os.path.exists(path) and os.path.getsize(path) == cls.TOKEN_BYTES):
                if time.monotonic() - t0 > timeout_s:
                    raise TimeoutError(f"NativeComm: no token at {path} after {timeout_s} s")
                time.sleep(0.05)
            with open(path, "rb") as f:
                return f.read()
        return exchange

    @classmethod
    def from_env(cls, path, timeout_s=300.0):
        """RANK / WORLD_SIZE from the environment (torchrun's names), token through ``file_exchange(path)``."""
        import os
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), cls.file_exchange(path, timeout_s))

    def _call(self, what, fn, t, *mid):
        if self._h is None:
            raise RuntimeError("NativeComm: used after close()")
        if not (t.is_cuda and t.is_contiguous()):
            raise ValueError("NativeComm: contiguous CUDA tensors only")
        f = self._ffi
        self._check(fn(self._h, f.ptr(t), t.numel(), f.dt(t), *mid, f.stream()), what)

    def all_reduce(self, t, average=False):
        """In place on the current stream: sum (or mean) over the ranks."""
        self._call("vvae_comm_allreduce", self._ffi.lib.vvae_comm_allreduce, t, 1 if average else 0)

    def broadcast(self, t, src=0):
        self._call("vvae_comm_broadcast", self._ffi.lib.vvae_comm_broadcast, t, int(src))

    def close(self):
        if self._h is not None:
            h, self._h = self._h, None
            self._check(self._ffi.lib.vvae_comm_destroy(h), "vvae_comm_destroy")

    def __del__(self):
        try:
            self.close()
        except Exception:          # interpreter shutdown: the library may already be gone
            pass


class FlatParams:
    """Re-homes every parameter (and its gradient) of ``model`` into one contiguous fp32 buffer each."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 7) // 8 * 8          # keep every fp32 AND bf16-shadow view 16-byte aligned
        self.total = off
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view(p.shape)
        self.shadow = None
        self._shadow_views = {}
        F_.invalidate_shadows()

    def broadcast(self, src=0, group=None, comm=None):
        """Replicate rank ``src``'s parameters on every rank (the reference loads on process 0 and calls
        ``broadcast_one_to_all``, claude_distributed/distributed_train.py:312-341): one collective on the flat buffer.
        ``comm``: a NativeComm to use instead of torch.distributed."""
        if comm is not None:
            if comm.world > 1:
                comm.broadcast(self.flat, src)
                self.refresh_shadow()
            return
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.flat, src=src, group=group)
            if self.shadow is not None:
                self.refresh_shadow()
            else:
                F_.params_changed()

    def all_reduce_grads(self, comm_dtype=torch.float32, group=None, comm=None):
        """Sum the flat gradient over the ranks with ONE collective (the graph-mode exchange step: the reference's XLA
        SPMD all-reduce of claude_distributed/distributed_train.py:378-380).  ``comm_dtype=torch.bfloat16`` sends a bf16
        copy (341 MB instead of 682 MB at production size): one cast kernel each way, the sum itself runs in NCCL; the
        fp32 buffer then holds the bf16-rounded sum -- every rank the SAME values, so replicas stay bit-identical.
        ``comm``: a NativeComm to use instead of torch.distributed (``vvae_comm_allreduce`` on the current stream)."""
        if comm is not None:
            if comm.world == 1:
                return
            if comm_dtype == torch.float32:
                comm.all_reduce(self.grad)
                return
            if getattr(self, "_grad_lowp", None) is None or self._grad_lowp.dtype != comm_dtype:
                self._grad_lowp = torch.empty(self.total, dtype=comm_dtype, device=self.grad.device)
            ops.cast_into(self.grad, self._grad_lowp)
            comm.all_reduce(self._grad_lowp)
            ops.cast_into(self._grad_lowp, self.grad)
            return
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if comm_dtype == torch.float32 or not self.grad.is_cuda:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return
        if getattr(self, "_grad_lowp", None) is None or self._grad_lowp.dtype != comm_dtype:
            self._grad_lowp = torch.empty(self.total, dtype=comm_dtype, device=self.grad.device)
        ops.cast_into(self.grad, self._grad_lowp)
        dist.all_reduce(self._grad_lowp, op=dist.ReduceOp.SUM, group=group)
        ops.cast_into(self._grad_lowp, self.grad)

    def zero_grad(self):
        if self.grad.is_cuda:
            ops.fill_(self.grad, 0.0)
        else:                      # host-side bookkeeping tests (gloo); no compute ever runs on CPU tensors
            self.grad.zero_()

    def enable_bf16_shadow(self):
        """Keep ONE flat bf16 copy of all parameters, refreshed by a single cast kernel per optimizer step."""
        self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=self.flat.device)
        for p, o in zip(self.params, self.offsets):
            self._shadow_views[id(p)] = (weakref.ref(p), self.shadow[o:o + p.numel()].view(p.shape))
        F_._flat_shadow_views.update(self._shadow_views)     # merge: several FlatParams may live in one process
        self.refresh_shadow()

    def refresh_shadow(self):
        if self.shadow is not None:
            ops.cast_into(self.flat, self.shadow)
        F_.params_changed()


class GradAllReducer:
    """Bucketed, backward-overlapped all-reduce (sum) of FlatParams.grad.

    A bucket is launched when every one of its parameters has reported ``uses`` times (default 1: each parameter is
    produced by exactly one backward Function per step, as in VideoVAE).  With weight sharing, a module called twice
    or micro-batch accumulation pass ``uses_per_step`` (a dict id(param) -> count, or an int for all) -- or call
    ``defer()`` to launch every bucket from ``finish_step`` only."""

    def __init__(self, flat: FlatParams, bucket_bytes=64 << 20, process_group=None, uses_per_step=1):
        self.flat = flat
        self.uses_per_step = uses_per_step
        self._deferred = False
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.stream = torch.cuda.Stream() if flat.flat.is_cuda else None
        # buckets are contiguous slices of the flat gradient, cut in REVERSE parameter order
        self.buckets = []      # (start, end, n_params)
        self.bucket_of = {}
        cur_end, cur_start, cur_n = flat.total, flat.total, 0
        for idx in range(len(flat.params) - 1, -1, -1):
            cur_start = flat.offsets[idx]
            cur_n += 1
            self.bucket_of[id(flat.params[idx])] = len(self.buckets)
            if (cur_end - cur_start) * 4 >= bucket_bytes or idx == 0:
                self.buckets.append((cur_start, cur_end, cur_n))
                cur_end, cur_n = cur_start, 0
        self._pending = None
        self._handles = []
        self._seen = None
        self.launch_order = []
        F_._grad_hooks.append(self._on_grads_ready)

    def defer(self, on=True):
        """Launch nothing during backward; ``finish_step`` reduces every bucket (safe for any gradient-use pattern)."""
        self._deferred = on

    def _uses(self, p):
        u = self.uses_per_step
        return u.get(id(p), 1) if isinstance(u, dict) else int(u)

    def start_step(self):
        self._pending = [n for (_, _, n) in self.buckets]
        self._seen = {}
        self._handles = []
        self.launch_order = []

    def _on_grads_ready(self, params):
        if self._pending is None or self.world == 1 or self._deferred:
            return
        for p in params:
            if p is None or id(p) not in self.bucket_of:
                continue
            n = self._seen.get(id(p), 0) + 1
            self._seen[id(p)] = n
            if n != self._uses(p):        # not yet its last use this step (or an unexpected extra one: see finish_step)
                continue
            b = self.bucket_of[id(p)]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)

    def _launch(self, b):
        s, e, _ = self.buckets[b]
        self.launch_order.append(b)
        if self.stream is None:    # CPU tensors (gloo tests of the bucketing logic): reduce in place, synchronously
            dist.all_reduce(self.flat.grad[s:e], op=dist.ReduceOp.SUM, group=self.pg)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self._handles.append(dist.all_reduce(self.flat.grad[s:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def finish_step(self):
        """Flush buckets that never filled (unused parameters) and join the communication stream."""
        if self.world == 1 or self._pending is None:
            return
        extra = [k for k, n in self._seen.items() if not self._deferred and n > self._uses_by_id(k)]
        if extra:
            raise RuntimeError(f"{len(extra)} parameter(s) received gradient more often than uses_per_step allows: their "
                               "bucket was all-reduced before the last accumulation (pass uses_per_step or call defer())")
        for b, n in enumerate(self._pending):
            if n > 0:
                self._pending[b] = 0
                self._launch(b)
        for h in self._handles:
            h.wait()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._pending = None

    def _uses_by_id(self, pid):
        u = self.uses_per_step
        return u.get(pid, 1) if isinstance(u, dict) else int(u)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self._on_grads_ready in F_._grad_hooks:
            F_._grad_hooks.remove(self._on_grads_ready)


class SplitAllReduce:
    """Gradient all-reduce for the graph-captured step: the decoder's slice of the flat gradient is reduced on a side
    stream as soon as the graph's ``decoder_done`` event fires (while the encoder's backward still runs inside the
    graph); the rest (encoder + fill token) follows the graph.  Parameters are laid out encoder | decoder | fill_token
    (``model.parameters()`` order), so both parts are contiguous slices."""

    def __init__(self, flat: FlatParams, model, process_group=None):
        self.flat, self.pg = flat, process_group
        ids = {id(p) for p in model.decoder.parameters()}
        idx = [i for i, p in enumerate(flat.params) if id(p) in ids]
        assert idx and idx == list(range(idx[0], idx[-1] + 1)), "decoder parameters must be contiguous in the flat buffer"
        self.lo = flat.offsets[idx[0]]
        self.hi = flat.offsets[idx[-1] + 1] if idx[-1] + 1 < len(flat.params) else flat.total
        self.stream = torch.cuda.Stream()

    def __call__(self, decoder_done_event):
        """Call right after graph.replay() was enqueued on the current stream."""
        g = self.flat.grad
        works = []
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(decoder_done_event)
            works.append(dist.all_reduce(g[self.lo:self.hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        if self.lo > 0:
            works.append(dist.all_reduce(g[:self.lo], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        if self.hi < self.flat.total:
            works.append(dist.all_reduce(g[self.hi:], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        for w in works:
            w.wait()
        torch.cuda.current_stream().wait_stream(self.stream)


class FlatAdam:
    """optax.chain(clip_by_global_norm(clip), adam(lr)) on the flat buffers (train/rl_nonadversarial.py:241-253):
    one reduction kernel + one update kernel per step, no host synchronisation."""

    def __init__(self, flat: FlatParams, lr=5e-5, b1=0.9, b2=0.999, eps=1e-8, clip=1.0):
        self.flat, self.lr, self.b1, self.b2, self.eps, self.clip = flat, lr, b1, b2, eps, clip
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=flat.flat.device)
        # fixed-order global-norm reduction: replicas that hold the same reduced gradient get the same clip scale bit for
        # bit and therefore stay identical (atomics would let them drift by an ulp per step)
        self._partials = (torch.empty(ops.sumsq_partials(flat.total), dtype=torch.float32, device=flat.flat.device)
                          if flat.flat.is_cuda else None)
        self.t = 0

    def step(self, grad_scale=1.0, lr=None):
        """``self.lr`` may be a float or a schedule ``count -> lr`` (count = 0 for the first update, as in optax)."""
        if lr is None:
            lr = self.lr(self.t) if callable(self.lr) else self.lr
        self.t += 1
        ops.fill_(self.gnorm_sq, 0.0)
        ops.sumsq_accum(self.flat.grad, self.gnorm_sq, self._partials)
        # the bf16 shadow of the parameters (what the bf16 kernels read) is written by the same pass
        ops.adam_step_(self.flat.flat, self.flat.grad, self.m, self.v, float(lr), self.b1, self.b2,
                       self.eps, self.t, self.gnorm_sq, self.clip, grad_scale, shadow=self.flat.shadow)
        if self.flat.shadow is not None:
            F_.params_changed()
        else:
            F_.invalidate_shadows()
