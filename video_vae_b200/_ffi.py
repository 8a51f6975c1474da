"""ctypes binding of libvvae.so (the C ABI declared in include/vvae.h).

There is no fallback of any kind: if the shared library is missing this module raises at import, and every
compute call raises ``VvaeError`` when the library reports a non-zero status (e.g. no CUDA device).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VVAE_LIB", os.path.join(_HERE, "libvvae.so"))

F32, BF16 = 0, 1
EPI_NONE, EPI_SILU, EPI_RESIDUAL, EPI_DSILU, EPI_QKNORM_ROPE = 0, 1, 2, 3, 4
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05 = 0, 1, 2


class VvaeError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python video_vae_b200/build.py` (nvcc, sm_100a). "
        "video_vae_b200 has no CPU or PyTorch fallback path.")

lib = C.CDLL(LIB_PATH)

vp, ll, i32, f32, u64 = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_ulonglong


class GemmArgs(C.Structure):
    _fields_ = [("M", i32), ("N", i32), ("K", i32),
                ("A", vp), ("lda", ll), ("transA", i32),
                ("B", vp), ("ldb", ll), ("transB", i32),
                ("C", vp), ("ldc", ll),
                ("dtype", i32), ("out_dtype", i32),
                ("bias", vp),
                ("epilogue", i32),
                ("aux_in", vp), ("ld_aux_in", ll),
                ("aux_out", vp), ("ld_aux_out", ll),
                ("accumulate", i32),
                ("backend", i32),
                ("bsum_accum", vp),
                ("qk_q_scale", vp), ("qk_k_scale", vp), ("rope_cos", vp), ("rope_sin", vp),
                ("rope_pos_div", ll), ("rope_pos_mod", i32), ("qk_heads", i32), ("qk_hd", i32), ("qk_eps", f32)]


class AttnArgs(C.Structure):
    _fields_ = [("n_outer", i32), ("n_inner", i32), ("L", i32), ("heads", i32), ("hd", i32),
                ("tok_stride_outer", ll), ("tok_stride_inner", ll), ("tok_stride_pos", ll),
                ("q", vp), ("k", vp), ("v", vp), ("o", vp),
                ("q_rs", ll), ("k_rs", ll), ("v_rs", ll), ("o_rs", ll),
                ("lse", vp),
                ("mask", vp), ("mask_seq_div", ll), ("ms_seq", ll), ("ms_head", ll), ("ms_q", ll), ("ms_k", ll),
                ("scale", f32),
                ("dtype", i32),
                ("d_o", vp), ("do_rs", ll),
                ("dq", vp), ("dk", vp), ("dv", vp), ("dq_rs", ll), ("dk_rs", ll), ("dv_rs", ll),
                ("delta", vp),
                ("backend", i32)]


class ConvArgs(C.Structure):
    _fields_ = [("B", i32), ("T", i32), ("H", i32), ("W", i32), ("Cin", i32), ("Cout", i32),
                ("kt", i32), ("kh", i32), ("kw", i32),
                ("x", vp), ("x_ld", ll),
                ("w", vp),
                ("bias", vp),
                ("y", vp), ("y_ld", ll),
                ("epilogue", i32), ("aux_in", vp), ("ld_aux", ll),
                ("dw_accum", vp),
                ("dtype", i32),
                ("backend", i32),
                ("wprep", vp),
                ("pad_out", i32)]


_SIGS = {
    "vvae_version": ([], i32),
    "vvae_device_ok": ([], i32),
    "vvae_debug_set": ([i32, ll], i32),
    "vvae_debug_get": ([i32, C.POINTER(u64)], i32),
    "vvae_cast": ([vp, i32, vp, i32, ll, vp], i32),
    "vvae_fill_f32": ([vp, f32, ll, vp], i32),
    "vvae_colsum": ([vp, ll, ll, i32, vp, i32, vp], i32),
    "vvae_gemm": ([C.POINTER(GemmArgs), vp], i32),
    "vvae_gemm_uses_tcgen05": ([C.POINTER(GemmArgs)], i32),
    "vvae_layernorm_fwd": ([vp, vp, vp, vp, vp, vp, ll, i32, f32, i32, vp], i32),
    "vvae_layernorm_bwd": ([vp, vp, vp, vp, vp, vp, vp, vp, vp, ll, i32, i32, vp], i32),
    "vvae_qknorm_rope_fwd": ([vp, vp, vp, vp, vp, vp, ll, i32, i32, ll, i32, f32, i32, vp], i32),
    "vvae_qknorm_rope_bwd": ([vp, vp, vp, vp, vp, vp, vp, vp, vp, ll, i32, i32, ll, i32, f32, i32, vp], i32),
    "vvae_attn_fwd": ([C.POINTER(AttnArgs), vp], i32),
    "vvae_attn_bwd": ([C.POINTER(AttnArgs), vp], i32),
    "vvae_patchify": ([vp, i32, vp, i32, i32, i32, i32, i32, i32, vp], i32),
    "vvae_pixel_shuffle": ([vp, vp, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    "vvae_pixel_shuffle_pitched": ([vp, vp, i32, i32, i32, i32, i32, i32, ll, i32, vp], i32),
    "vvae_conv3d_fwd": ([C.POINTER(ConvArgs), vp], i32),
    "vvae_conv3d_dgrad": ([C.POINTER(ConvArgs), vp], i32),
    "vvae_conv3d_wgrad": ([C.POINTER(ConvArgs), vp], i32),
    "vvae_conv3d_wprep_bytes": ([C.POINTER(ConvArgs), i32], ll),
    "vvae_conv3d_wprep": ([C.POINTER(ConvArgs), i32, vp, vp], i32),
    "vvae_convT122_workspace_bytes": ([i32, i32, i32, i32, i32], ll),
    "vvae_convT122_fwd": ([vp, vp, vp, vp, ll, i32, i32, i32, i32, i32, i32, vp, ll, vp], i32),
    "vvae_convT122_bwd": ([vp, ll, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, ll, vp], i32),
    "vvae_groupnorm_silu_fwd": ([vp, vp, ll, vp, vp, vp, vp, vp, i32, ll, i32, i32, f32, i32, vp], i32),
    "vvae_groupnorm_silu_bwd": ([vp, ll, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, ll, i32, i32, i32, vp], i32),
    "vvae_maxpool122_fwd": ([vp, ll, vp, i32, i32, i32, i32, i32, vp], i32),
    "vvae_maxpool122_bwd": ([vp, ll, vp, vp, ll, vp, i32, i32, i32, i32, i32, vp], i32),
    "vvae_copy_channels": ([vp, ll, ll, vp, ll, ll, ll, i32, i32, vp], i32),
    "vvae_softplus_log_fwd": ([vp, vp, ll, i32, vp], i32),
    "vvae_softplus_log_bwd": ([vp, vp, vp, ll, i32, vp], i32),
    "vvae_selection_fwd": ([vp, vp, vp, vp, u64, u64, i32, f32, vp, vp, vp, i32, i32, i32, vp], i32),
    "vvae_reparam_gate_fwd": ([vp, vp, vp, u64, u64, vp, vp, vp, vp, vp, ll, i32, i32, i32, i32, vp], i32),
    "vvae_reparam_gate_bwd": ([vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ll, i32, i32, i32, i32, vp], i32),
    "vvae_recon_loss_fwd": ([vp, i32, vp, vp, vp, vp, i32, i32, ll, i32, vp], i32),
    "vvae_recon_loss_bwd": ([vp, i32, vp, vp, vp, f32, f32, f32, vp, vp, i32, i32, ll, i32, vp], i32),
    "vvae_kl_fwd": ([vp, vp, vp, vp, ll, i32, i32, i32, vp], i32),
    "vvae_relu_fwd": ([vp, vp, ll, i32, vp], i32),
    "vvae_relu_bwd": ([vp, vp, vp, ll, i32, vp], i32),
    "vvae_vgg_preprocess_fwd": ([vp, i32, vp, ll, i32, i32, vp], i32),
    "vvae_vgg_preprocess_bwd": ([vp, vp, ll, i32, i32, vp], i32),
    "vvae_recon_loss_per_sample_fwd": ([vp, i32, vp, vp, vp, vp, i32, i32, ll, i32, vp], i32),
    "vvae_kl_per_sample_fwd": ([vp, vp, vp, vp, i32, ll, i32, i32, i32, vp], i32),
    "vvae_kl_bwd": ([vp, vp, vp, f32, vp, vp, vp, ll, i32, i32, i32, vp], i32),
    "vvae_philox_fill": ([vp, ll, u64, u64, i32, vp], i32),
    "vvae_sumsq_f32": ([vp, ll, vp, vp], i32),
    "vvae_sumsq_partials": ([ll], i32),
    "vvae_sumsq_f32_det": ([vp, ll, vp, vp, vp], i32),
    "vvae_adam_step": ([vp, vp, vp, vp, vp, ll, f32, f32, f32, f32, i32, vp, f32, f32, vp], i32),
    # gradient exchange behind the C ABI (NCCL resolved with dlopen; ddp.NativeComm)
    "vvae_workspace_bytes": ([i32, C.POINTER(ll), i32], ll),
    "vvae_comm_unique_id": ([vp], i32),
    "vvae_comm_init": ([C.POINTER(vp), vp, i32, i32], i32),
    "vvae_comm_rank": ([vp, C.POINTER(i32), C.POINTER(i32)], i32),
    "vvae_comm_allreduce": ([vp, vp, ll, i32, i32, vp], i32),
    "vvae_comm_broadcast": ([vp, vp, ll, i32, i32, vp], i32),
    "vvae_comm_destroy": ([vp], i32),
}

EXPORTED = tuple(_SIGS) + ("vvae_last_error",)

for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = _res
lib.vvae_last_error.argtypes = []
lib.vvae_last_error.restype = C.c_char_p

launch_count = 0  # number of libvvae compute entry points invoked (bench.py reports it as gpu_launches evidence)


def check(rc, what=""):
    global launch_count
    launch_count += 1
    if rc != 0:
        raise VvaeError(f"{what} failed (status {rc}): {lib.vvae_last_error().decode()}")


def dt(t_or_dtype):
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise VvaeError(f"unsupported dtype {d}: libvvae computes in float32 or bfloat16")


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise VvaeError("libvvae operates on CUDA tensors only (there is no CPU path)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_device():
    if not torch.cuda.is_available() or not lib.vvae_device_ok():
        raise VvaeError("video_vae_b200 needs a CUDA device of compute capability 10.x (B200); there is no CPU path")


# ---------------------------------------------------------------------------------------------------- profiling aid
class AbiProfile:
    """CUDA-event timing of EVERY libvvae entry point that takes a stream (bench.py --profile-kernels, scripts/):

        with AbiProfile() as prof: step()
        torch.cuda.synchronize(); table = prof.table()

    Each call is bracketed by two events on the current stream, so the numbers are device times of that entry point's
    kernels in an eager pass (launch gaps included in neither).  Not for use under CUDA-graph capture."""

    _NO_STREAM = ("vvae_version", "vvae_device_ok", "vvae_debug_set", "vvae_debug_get", "vvae_gemm_uses_tcgen05", "vvae_conv3d_wprep_bytes",
                  "vvae_convT122_workspace_bytes", "vvae_sumsq_partials", "vvae_workspace_bytes", "vvae_comm_unique_id",
                  "vvae_comm_init", "vvae_comm_rank", "vvae_comm_destroy")

    def __init__(self, key=None):
        self.records, self._saved, self.key = [], {}, key

    def __enter__(self):
        for name in _SIGS:
            if name in self._NO_STREAM:
                continue
            fn = getattr(lib, name)
            self._saved[name] = fn

            def wrapper(*args, _fn=fn, _name=name):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = _fn(*args)
                e1.record()
                self.records.append((self.key(_name, args) if self.key else _name, e0, e1))
                return rc
            setattr(lib, name, wrapper)
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(lib, name, fn)
        return False

    def table(self):
        """[(entry point, calls, total ms)] sorted by time; call after torch.cuda.synchronize()."""
        agg = {}
        for name, e0, e1 in self.records:
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
        return sorted(((n, c, t) for n, (c, t) in agg.items()), key=lambda r: -r[2])
