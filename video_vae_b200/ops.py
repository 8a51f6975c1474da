"""Thin tensor-level wrappers over the libvvae C ABI (include/vvae.h).

Every function enqueues one (or a few) hand-written CUDA kernels on torch's current stream.  PyTorch supplies device
memory and the stream only.  No function here has a CPU or PyTorch-op fallback.
"""
import ctypes as C

import torch

from ._ffi import (EPI_QKNORM_ROPE, BACKEND_AUTO, BACKEND_SIMT, BF16, EPI_NONE, EPI_RESIDUAL, AttnArgs, ConvArgs, GemmArgs,
                   check, dt, lib, ptr, stream)

LN_EPS = 1e-6

CONV_BACKEND = BACKEND_AUTO  # tests force BACKEND_SIMT to cross-check the tensor-core conv against the generic one
ATTN_BACKEND = BACKEND_AUTO  # same for attention

# bench.py sets PROFILE to a list; instrumented ops then append (class, flops, bytes, start_event, stop_event).
PROFILE = None


class _Prof:
    __slots__ = ("name", "flops", "nbytes", "e0")

    def __init__(self, name, flops, nbytes=0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.name, self.flops, self.nbytes, self.e0, e1))
        return False


def empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


def cast(src, dtype):
    """dst = src.to(dtype) with libvvae's cast kernel (fp32 <-> bf16)."""
    if src.dtype == dtype:
        return src
    src = src.contiguous()
    dst = torch.empty_like(src, dtype=dtype)
    check(lib.vvae_cast(ptr(src), dt(src), ptr(dst), dt(dtype), src.numel(), stream()), "vvae_cast")
    return dst


def cast_into(src, dst):
    check(lib.vvae_cast(ptr(src), dt(src), ptr(dst), dt(dst), src.numel(), stream()), "vvae_cast")
    return dst


def fill_(t, value=0.0):
    assert t.dtype == torch.float32 and t.is_contiguous()
    check(lib.vvae_fill_f32(ptr(t), float(value), t.numel(), stream()), "vvae_fill_f32")
    return t


def zeros_f32(shape, device):
    return fill_(torch.empty(shape, dtype=torch.float32, device=device))


def colsum_accum(x2d, out):
    """out[n] += sum_rows x2d[:, n]   (x2d may be a column slice: stride(1) == 1)."""
    assert x2d.dim() == 2 and x2d.stride(1) == 1 and out.dtype == torch.float32
    check(lib.vvae_colsum(ptr(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], ptr(out), dt(x2d), stream()), "vvae_colsum")


def gemm(A, B, *, transA=False, transB=False, out=None, out_dtype=None, bias=None, epilogue=EPI_NONE, aux_in=None,
         aux_out=None, accumulate=False, backend=BACKEND_AUTO, bsum=None, qknorm=None):
    """C = epilogue(op(A) @ op(B)); A, B 2-D with unit inner stride (views with a row stride are fine).
    bsum (fp32 [N], optional) += column sums of B: the bias gradient of a Linear, fused into its weight-gradient GEMM.
    qknorm = (q_scale, k_scale, cos, sin, heads, hd, pos_div, pos_mod) with epilogue=EPI_QKNORM_ROPE: the QKV
    projection; aux_out [M, 2*heads*hd] receives rope(LN(q)) | rope(LN(k))."""
    assert A.dim() == 2 and B.dim() == 2 and A.stride(1) == 1 and B.stride(1) == 1 and A.dtype == B.dtype
    M, K = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
    Kb, N = (B.shape[1], B.shape[0]) if transB else (B.shape[0], B.shape[1])
    assert K == Kb, f"gemm: contraction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or A.dtype, device=A.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    a = GemmArgs(M, N, K, ptr(A), A.stride(0), int(transA), ptr(B), B.stride(0), int(transB), ptr(out), out.stride(0),
                 dt(A), dt(out), ptr(bias), epilogue,
                 ptr(aux_in), aux_in.stride(0) if aux_in is not None else 0,
                 ptr(aux_out), aux_out.stride(0) if aux_out is not None else 0,
                 int(accumulate), backend, ptr(bsum))
    if qknorm is not None:
        qs, ks, cos, sin, heads, hd, pos_div, pos_mod = qknorm
        a.qk_q_scale, a.qk_k_scale, a.rope_cos, a.rope_sin = ptr(qs), ptr(ks), ptr(cos), ptr(sin)
        a.rope_pos_div, a.rope_pos_mod, a.qk_heads, a.qk_hd, a.qk_eps = pos_div, pos_mod, heads, hd, LN_EPS
    if PROFILE is not None:
        name = "gemm_tcgen05" if lib.vvae_gemm_uses_tcgen05(C.byref(a)) else "gemm_simt"
        with _Prof(name, 2.0 * M * N * K):
            check(lib.vvae_gemm(C.byref(a), stream()), "vvae_gemm")
        return out
    check(lib.vvae_gemm(C.byref(a), stream()), "vvae_gemm")
    return out


# ---------------------------------------------------------------- norms
def layernorm_fwd(x2d, gamma, beta, save_stats=True):
    rows, D = x2d.shape
    y = torch.empty_like(x2d)
    mean = torch.empty(rows, dtype=torch.float32, device=x2d.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x2d.device) if save_stats else None
    check(lib.vvae_layernorm_fwd(ptr(x2d), ptr(y), ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), rows, D, LN_EPS, dt(x2d),
                                 stream()), "vvae_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x2d, mean, rstd, gamma, dres, dgamma, dbeta, out=None):
    rows, D = x2d.shape
    dx = torch.empty_like(x2d) if out is None else out
    check(lib.vvae_layernorm_bwd(ptr(dy), ptr(x2d), ptr(mean), ptr(rstd), ptr(gamma), ptr(dres), ptr(dx), ptr(dgamma),
                                 ptr(dbeta), rows, D, dt(x2d), stream()), "vvae_layernorm_bwd")
    return dx


def qkv_projection(h, w, bias, q_scale, k_scale, cos, sin, heads, hd, pos_div, pos_mod):
    """qkv = h @ w + bias  and  qk = rope(LN(q)) | rope(LN(k))  in ONE call (train/layers.py:160-166)."""
    rows = h.shape[0]
    qkv = torch.empty((rows, w.shape[1]), dtype=h.dtype, device=h.device)
    qk = torch.empty((rows, 2 * heads * hd), dtype=h.dtype, device=h.device)
    gemm(h, w, out=qkv, bias=bias, epilogue=EPI_QKNORM_ROPE, aux_out=qk,
         qknorm=(q_scale, k_scale, cos, sin, heads, hd, pos_div, pos_mod))
    return qkv, qk


def qknorm_rope_fwd(qkv, q_scale, k_scale, cos, sin, heads, hd, pos_div, pos_mod):
    rows = qkv.shape[0]
    out = torch.empty((rows, 2 * heads * hd), dtype=qkv.dtype, device=qkv.device)
    check(lib.vvae_qknorm_rope_fwd(ptr(qkv), ptr(out), ptr(q_scale), ptr(k_scale), ptr(cos), ptr(sin), rows, heads, hd,
                                   pos_div, pos_mod, LN_EPS, dt(qkv), stream()), "vvae_qknorm_rope_fwd")
    return out


def qknorm_rope_bwd_(dqkv, qkv, q_scale, k_scale, cos, sin, dq_scale, dk_scale, heads, hd, pos_div, pos_mod, dbias_qk=None):
    """In place on the q|k part of dqkv.  dbias_qk (fp32 [2*heads*hd], optional) += column sums of the result."""
    rows = qkv.shape[0]
    check(lib.vvae_qknorm_rope_bwd(ptr(dqkv), ptr(qkv), ptr(q_scale), ptr(k_scale), ptr(cos), ptr(sin), ptr(dq_scale),
                                   ptr(dk_scale), ptr(dbias_qk), rows, heads, hd, pos_div, pos_mod, LN_EPS, dt(qkv), stream()),
          "vvae_qknorm_rope_bwd")
    return dqkv


# ---------------------------------------------------------------- attention
class AttnGeom:
    """Which tokens form a sequence (see include/vvae.h, vvae_attn_args)."""

    def __init__(self, n_outer, n_inner, L, ts_outer, ts_inner, ts_pos):
        self.n_outer, self.n_inner, self.L = n_outer, n_inner, L
        self.ts = (ts_outer, ts_inner, ts_pos)

    @property
    def n_seq(self):
        return self.n_outer * self.n_inner


class AttnMask:
    """uint8 mask plus the strides that broadcast it to (sequence, head, query, key)."""

    def __init__(self, data, seq_div, ms_seq, ms_head, ms_q, ms_k):
        self.data, self.seq_div, self.strides = data, seq_div, (ms_seq, ms_head, ms_q, ms_k)


def _attn_args(geom, heads, hd, q, k, v, o, lse, mask, scale):
    a = AttnArgs()
    a.n_outer, a.n_inner, a.L, a.heads, a.hd = geom.n_outer, geom.n_inner, geom.L, heads, hd
    a.tok_stride_outer, a.tok_stride_inner, a.tok_stride_pos = geom.ts
    a.q, a.k, a.v, a.o = ptr(q), ptr(k), ptr(v), ptr(o)
    a.q_rs, a.k_rs, a.v_rs, a.o_rs = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    a.lse = ptr(lse)
    if mask is not None:
        a.mask = ptr(mask.data)
        a.mask_seq_div = mask.seq_div
        a.ms_seq, a.ms_head, a.ms_q, a.ms_k = mask.strides
    else:
        a.mask, a.mask_seq_div = None, 1
    a.scale = scale
    a.dtype = dt(q)
    a.backend = ATTN_BACKEND
    return a


def attn_fwd(geom, heads, hd, q, k, v, mask, scale):
    """q,k,v: 2-D token-major views [tokens, >= heads*hd] (row stride arbitrary). Returns o [tokens, heads*hd], lse."""
    n_tok = q.shape[0]
    o = torch.empty((n_tok, heads * hd), dtype=q.dtype, device=q.device)
    lse = torch.empty((geom.n_seq, heads, geom.L), dtype=torch.float32, device=q.device)
    a = _attn_args(geom, heads, hd, q, k, v, o, lse, mask, scale)
    with _Prof("attn_fwd_L%d" % geom.L, 4.0 * geom.n_seq * heads * geom.L * geom.L * hd):
        check(lib.vvae_attn_fwd(C.byref(a), stream()), "vvae_attn_fwd")
    return o, lse


def attn_bwd(geom, heads, hd, q, k, v, o, lse, d_o, dq, dk, dv, mask, scale):
    a = _attn_args(geom, heads, hd, q, k, v, o, lse, mask, scale)
    delta = torch.empty_like(lse)
    a.d_o, a.do_rs = ptr(d_o), d_o.stride(0)
    a.dq, a.dk, a.dv = ptr(dq), ptr(dk), ptr(dv)
    a.dq_rs, a.dk_rs, a.dv_rs = dq.stride(0), dk.stride(0), dv.stride(0)
    a.delta = ptr(delta)
    with _Prof("attn_bwd_L%d" % geom.L, 10.0 * geom.n_seq * heads * geom.L * geom.L * hd):
        check(lib.vvae_attn_bwd(C.byref(a), stream()), "vvae_attn_bwd")


# ---------------------------------------------------------------- rearrangements
def patchify(video, P, dtype):
    b, t, H, W, Cc = video.shape
    video = video.contiguous()
    out = torch.empty((b, t, (H // P) * (W // P), P * P * Cc), dtype=dtype, device=video.device)
    check(lib.vvae_patchify(ptr(video), dt(video), ptr(out), b * t, H, W, Cc, P, dt(dtype), stream()), "vvae_patchify")
    return out


def pixel_shuffle(src, b_t, H, W, CU, P, to_tokens, vox_ld=None):
    """``vox_ld`` (> CU): the voxel side has that many elements per pixel; to_tokens=False returns the padded buffer
    [b_t, H, W, vox_ld] with zeroed pad channels."""
    vox_ld = vox_ld or CU
    if to_tokens:
        dst = torch.empty((b_t * (H // P) * (W // P), P * P * CU), dtype=src.dtype, device=src.device)
    else:
        dst = torch.empty((b_t, H, W, vox_ld), dtype=src.dtype, device=src.device)
    if vox_ld != CU:
        check(lib.vvae_pixel_shuffle_pitched(ptr(src), ptr(dst), b_t, H, W, CU, P, int(to_tokens), vox_ld, dt(src), stream()),
              "vvae_pixel_shuffle_pitched")
        return dst
    check(lib.vvae_pixel_shuffle(ptr(src), ptr(dst), b_t, H, W, CU, P, int(to_tokens), dt(src), stream()),
          "vvae_pixel_shuffle")
    return dst


# ---------------------------------------------------------------- convolutions (channels-last, 5-D [B,T,H,W,C])
def _conv_args(x, x_ld, w, bias, y, y_ld, B, T, H, W, Cin, Cout, ks, epilogue=EPI_NONE, aux=None, aux_ld=0, dw=None,
               wprep=None, dtype=None, pad_out=False):
    a = ConvArgs()
    a.B, a.T, a.H, a.W, a.Cin, a.Cout = B, T, H, W, Cin, Cout
    a.kt, a.kh, a.kw = ks
    a.x, a.x_ld, a.w, a.bias, a.y, a.y_ld = ptr(x), x_ld, ptr(w), ptr(bias), ptr(y), y_ld
    a.epilogue, a.aux_in, a.ld_aux = epilogue, ptr(aux), aux_ld
    a.dw_accum = ptr(dw)
    a.dtype = dtype if dtype is not None else dt(x if x is not None else y)
    a.backend = CONV_BACKEND
    a.wprep = ptr(wprep)
    a.pad_out = int(pad_out)
    return a


def conv3d_wprep(w, which, B, T, H, W, Cin, Cout, ks, x_ld, y_ld):
    """Weight image for the tensor-core conv path (None when the shape runs on the generic kernel)."""
    if w.dtype != torch.bfloat16:
        return None
    a = _conv_args(None, x_ld, w, None, None, y_ld, B, T, H, W, Cin, Cout, ks, dtype=BF16)
    nbytes = lib.vvae_conv3d_wprep_bytes(C.byref(a), which)
    if nbytes <= 0:
        return None
    img = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    check(lib.vvae_conv3d_wprep(C.byref(a), which, ptr(img), stream()), "vvae_conv3d_wprep")
    return img


def conv3d_fwd(x, w, bias, ks, Cin, Cout, x_ld=None, residual=None, out=None, out_ld=None, wprep=None, pad_out=False):
    """x: [B,T,H,W,x_ld] storage whose first Cin channels are the input; w: [kt,kh,kw,Cin,Cout] (compute dtype).
    ``out``/``out_ld`` let the result land in (a channel slice of) a wider pre-allocated buffer."""
    B, T, H, W = x.shape[:4]
    if out is None:
        out = torch.empty((B, T, H, W, Cout), dtype=x.dtype, device=x.device)
        out_ld = Cout
    a = _conv_args(x, x_ld or x.shape[-1], w, bias, out, out_ld, B, T, H, W, Cin, Cout, ks,
                   EPI_RESIDUAL if residual is not None else EPI_NONE, residual, Cout, wprep=wprep, pad_out=pad_out)
    with _Prof("conv3d_fwd", 2.0 * B * T * H * W * ks[0] * ks[1] * ks[2] * Cin * Cout):
        check(lib.vvae_conv3d_fwd(C.byref(a), stream()), "vvae_conv3d_fwd")
    return out


def conv3d_dgrad(dy, w, ks, Cin, Cout, dy_ld=None, out=None, out_ld=None, wprep=None, pad_out=False):
    B, T, H, W = dy.shape[:4]
    if out is None:
        out = torch.empty((B, T, H, W, Cin), dtype=dy.dtype, device=dy.device)
        out_ld = Cin
    a = _conv_args(out, out_ld, w, None, dy, dy_ld or dy.shape[-1], B, T, H, W, Cin, Cout, ks, wprep=wprep,
                   pad_out=pad_out)
    with _Prof("conv3d_dgrad", 2.0 * B * T * H * W * ks[0] * ks[1] * ks[2] * Cin * Cout):
        check(lib.vvae_conv3d_dgrad(C.byref(a), stream()), "vvae_conv3d_dgrad")
    return out


def conv3d_wgrad_accum(x, dy, dw, ks, Cin, Cout, x_ld=None, dy_ld=None):
    B, T, H, W = x.shape[:4]
    a = _conv_args(x, x_ld or x.shape[-1], None, None, dy, dy_ld or dy.shape[-1], B, T, H, W, Cin, Cout, ks, dw=dw)
    with _Prof("conv3d_wgrad", 2.0 * B * T * H * W * ks[0] * ks[1] * ks[2] * Cin * Cout):
        check(lib.vvae_conv3d_wgrad(C.byref(a), stream()), "vvae_conv3d_wgrad")


def _convt_workspace(b_t, H, W, Cin, Cout, device):
    """Scratch for the tensor-core ConvTranspose path (re-laid-out weights, dense GEMM-side image of y / dy)."""
    if CONV_BACKEND == BACKEND_SIMT:
        return None                       # generic kernel requested (tests compare the two paths)
    return torch.empty(int(lib.vvae_convT122_workspace_bytes(b_t, H, W, Cin, Cout)), dtype=torch.uint8, device=device)


def convT122_fwd(x, w, bias, Cout, out, out_ld):
    """x [B,T,H,W,Cin] -> writes out[..., :Cout] of a [B,T,2H,2W,out_ld] buffer."""
    B, T, H, W, Cin = x.shape
    ws = _convt_workspace(B * T, H, W, Cin, Cout, x.device)
    check(lib.vvae_convT122_fwd(ptr(x), ptr(w), ptr(bias), ptr(out), out_ld, B * T, H, W, Cin, Cout, dt(x), ptr(ws),
                                ws.numel() if ws is not None else 0, stream()), "vvae_convT122_fwd")


def convT122_bwd(dy, dy_ld, x, w, dw, Cout):
    B, T, H, W, Cin = x.shape
    dx = torch.empty_like(x)
    ws = _convt_workspace(B * T, H, W, Cin, Cout, x.device)
    check(lib.vvae_convT122_bwd(ptr(dy), dy_ld, ptr(x), ptr(w), ptr(dx), ptr(dw), B * T, H, W, Cin, Cout, dt(x), ptr(ws),
                                ws.numel() if ws is not None else 0, stream()), "vvae_convT122_bwd")
    return dx


def groupnorm_silu_fwd(x, gamma, beta, G, out=None, out_ld=None):
    B, Cc = x.shape[0], x.shape[-1]
    S = x.numel() // (B * Cc)
    if out is None:
        out, out_ld = torch.empty_like(x), Cc
    mean = torch.empty((B, G), dtype=torch.float32, device=x.device)
    rstd = torch.empty((B, G), dtype=torch.float32, device=x.device)
    stats = torch.empty((B, G, 2), dtype=torch.float32, device=x.device)
    check(lib.vvae_groupnorm_silu_fwd(ptr(x), ptr(out), out_ld, ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(stats),
                                      B, S, Cc, G, LN_EPS, dt(x), stream()), "vvae_groupnorm_silu_fwd")
    return out, mean, rstd


def groupnorm_silu_bwd(dy, dy_ld, x, gamma, beta, mean, rstd, dgamma, dbeta, G, dx_colsum=None):
    """dx_colsum (fp32 [C], optional) += per-channel sums of the returned dx (the preceding conv's bias gradient)."""
    B, Cc = x.shape[0], x.shape[-1]
    S = x.numel() // (B * Cc)
    dx = torch.empty_like(x)
    stats = torch.empty((B, G, 2), dtype=torch.float32, device=x.device)
    check(lib.vvae_groupnorm_silu_bwd(ptr(dy), dy_ld, ptr(x), ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(dx),
                                      ptr(dgamma), ptr(dbeta), ptr(stats), ptr(dx_colsum), B, S, Cc, G, dt(x), stream()),
          "vvae_groupnorm_silu_bwd")
    return dx


def maxpool122_fwd(x, x_ld, Cc):
    B, T, H, W = x.shape[:4]
    y = torch.empty((B, T, H // 2, W // 2, Cc), dtype=x.dtype, device=x.device)
    check(lib.vvae_maxpool122_fwd(ptr(x), x_ld, ptr(y), B * T, H, W, Cc, dt(x), stream()), "vvae_maxpool122_fwd")
    return y


def maxpool122_bwd(x, x_ld, dy, dskip, dskip_ld, Cc):
    B, T, H, W = x.shape[:4]
    dx = torch.empty((B, T, H, W, Cc), dtype=x.dtype, device=x.device)
    check(lib.vvae_maxpool122_bwd(ptr(x), x_ld, ptr(dy), ptr(dskip), dskip_ld, ptr(dx), B * T, H, W, Cc, dt(x), stream()),
          "vvae_maxpool122_bwd")
    return dx


def copy_channels(src, src_ld, src_off, dst, dst_ld, dst_off, rows, Cc):
    check(lib.vvae_copy_channels(ptr(src), src_ld, src_off, ptr(dst), dst_ld, dst_off, rows, Cc, dt(src), stream()),
          "vvae_copy_channels")


# ---------------------------------------------------------------- latent head / losses
def softplus_log_fwd(a):
    lv = torch.empty_like(a)
    check(lib.vvae_softplus_log_fwd(ptr(a), ptr(lv), a.numel(), dt(a), stream()), "vvae_softplus_log_fwd")
    return lv


def softplus_log_bwd(dlv, a):
    da = torch.empty_like(a)
    check(lib.vvae_softplus_log_bwd(ptr(dlv), ptr(a), ptr(da), a.numel(), dt(a), stream()), "vvae_softplus_log_bwd")
    return da


def selection_fwd(s1, w2, b2, u, seed, offset, train, temperature, bt, hw):
    dev = s1.device
    logit = torch.empty(bt, dtype=torch.float32, device=dev)
    p = torch.empty(bt, dtype=torch.float32, device=dev)
    sel = torch.empty(bt, dtype=torch.float32, device=dev)
    check(lib.vvae_selection_fwd(ptr(s1), ptr(w2), ptr(b2), ptr(u), seed, offset, int(train), float(temperature),
                                 ptr(logit), ptr(p), ptr(sel), bt, hw, dt(s1), stream()), "vvae_selection_fwd")
    return logit, p, sel


def reparam_gate_fwd(mean, logvar, eps, seed, offset, sel, fill, tok_per_frame, train, want_lowp):
    Dl = mean.shape[-1]
    n_tok = mean.numel() // Dl
    dev = mean.device
    c32 = torch.empty(mean.shape, dtype=torch.float32, device=dev)
    cT = torch.empty_like(mean) if want_lowp else None
    eps_out = None
    if train and eps is None:
        eps_out = torch.empty(mean.shape, dtype=torch.float32, device=dev)
    check(lib.vvae_reparam_gate_fwd(ptr(mean), ptr(logvar), ptr(eps), seed, offset, ptr(eps_out), ptr(sel), ptr(fill),
                                    ptr(c32), ptr(cT), n_tok, tok_per_frame, Dl, int(train), dt(mean), stream()),
          "vvae_reparam_gate_fwd")
    return c32, cT, (eps if eps is not None else eps_out)


def reparam_gate_bwd(dc, mean, logvar, eps, sel, fill, dmean_in, dlogvar_in, dfill, dsel, tok_per_frame, train):
    Dl = mean.shape[-1]
    n_tok = mean.numel() // Dl
    dmean, dlogvar = torch.empty_like(mean), torch.empty_like(logvar)
    check(lib.vvae_reparam_gate_bwd(ptr(dc), ptr(mean), ptr(logvar), ptr(eps), ptr(sel), ptr(fill), ptr(dmean_in),
                                    ptr(dlogvar_in), ptr(dmean), ptr(dlogvar), ptr(dfill), ptr(dsel), n_tok,
                                    tok_per_frame, Dl, int(train), dt(mean), stream()), "vvae_reparam_gate_bwd")
    return dmean, dlogvar


def recon_loss_fwd(video, recon, frame_mask, inv_len, out2):
    B, T = video.shape[:2]
    per_frame = video.numel() // (B * T)
    check(lib.vvae_recon_loss_fwd(ptr(video), dt(video), ptr(recon), ptr(frame_mask), ptr(inv_len), ptr(out2), B, T,
                                  per_frame, dt(recon), stream()), "vvae_recon_loss_fwd")


def recon_loss_per_sample_fwd(video, recon, frame_mask, inv_len, out_b2):
    """out_b2 fp32 [B,2] += per-sample (squared, absolute) masked error sums (RL loss, rl_nonadversarial.py:114-121)."""
    B, T = video.shape[:2]
    per_frame = video.numel() // (B * T)
    check(lib.vvae_recon_loss_per_sample_fwd(ptr(video), dt(video), ptr(recon), ptr(frame_mask), ptr(inv_len),
                                             ptr(out_b2), B, T, per_frame, dt(recon), stream()),
          "vvae_recon_loss_per_sample_fwd")


def recon_loss_bwd(video, recon, frame_mask, inv_len, w_mse, w_mae, inv_count, gscale=None):
    """gscale: optional fp32 device scalar multiplied in (the upstream gradient), so no host read is needed."""
    B, T = video.shape[:2]
    per_frame = video.numel() // (B * T)
    d = torch.empty_like(recon)
    check(lib.vvae_recon_loss_bwd(ptr(video), dt(video), ptr(recon), ptr(frame_mask), ptr(inv_len), float(w_mse),
                                  float(w_mae), float(inv_count), ptr(gscale), ptr(d), B, T, per_frame, dt(recon), stream()),
          "vvae_recon_loss_bwd")
    return d


def kl_fwd(mean, logvar, frame_w, out1, tok_per_frame):
    Dl = mean.shape[-1]
    check(lib.vvae_kl_fwd(ptr(mean), ptr(logvar), ptr(frame_w), ptr(out1), mean.numel() // Dl, tok_per_frame, Dl,
                          dt(mean), stream()), "vvae_kl_fwd")


def kl_per_sample_fwd(mean, logvar, frame_w, out_b, tok_per_frame):
    """out_b fp32 [B] += per-sample KL sums (rl_nonadversarial.py:145-146)."""
    B, Dl = mean.shape[0], mean.shape[-1]
    check(lib.vvae_kl_per_sample_fwd(ptr(mean), ptr(logvar), ptr(frame_w), ptr(out_b), B, mean.numel() // (Dl * B),
                                     tok_per_frame, Dl, dt(mean), stream()), "vvae_kl_per_sample_fwd")


def kl_bwd(mean, logvar, frame_w, scale, tok_per_frame, gscale=None):
    Dl = mean.shape[-1]
    dmean, dlogvar = torch.empty_like(mean), torch.empty_like(logvar)
    check(lib.vvae_kl_bwd(ptr(mean), ptr(logvar), ptr(frame_w), float(scale), ptr(gscale), ptr(dmean), ptr(dlogvar),
                          mean.numel() // Dl, tok_per_frame, Dl, dt(mean), stream()), "vvae_kl_bwd")
    return dmean, dlogvar


def relu_(x):
    """In-place ReLU (VGG feature extractor, train/vgg_tests.py)."""
    check(lib.vvae_relu_fwd(ptr(x), ptr(x), x.numel(), dt(x), stream()), "vvae_relu_fwd")
    return x


def relu_bwd(dy, y):
    dx = torch.empty_like(y)
    check(lib.vvae_relu_bwd(ptr(dy), ptr(y), ptr(dx), y.numel(), dt(y), stream()), "vvae_relu_bwd")
    return dx


def vgg_preprocess_fwd(x, dtype, ld=16):
    """x [b,t,H,W,3] in [0,1] -> ImageNet-normalised [b,t,H,W,ld] in ``dtype`` with zero pad channels."""
    x = x.contiguous()
    y = torch.empty(tuple(x.shape[:4]) + (ld,), dtype=dtype, device=x.device)
    check(lib.vvae_vgg_preprocess_fwd(ptr(x), dt(x), ptr(y), x.numel() // 3, ld, dt(y), stream()),
          "vvae_vgg_preprocess_fwd")
    return y


def vgg_preprocess_bwd(dy):
    dx = torch.empty(tuple(dy.shape[:4]) + (3,), dtype=dy.dtype, device=dy.device)
    check(lib.vvae_vgg_preprocess_bwd(ptr(dy), ptr(dx), dy.numel() // dy.shape[-1], dy.shape[-1], dt(dy), stream()),
          "vvae_vgg_preprocess_bwd")
    return dx


def philox_fill_(out, seed, offset, kind):
    """kind 'normal' = the reparameterisation draws, 'uniform' = the Gumbel-gate draws of the same (seed, offset)."""
    assert out.dtype == torch.float32 and out.is_contiguous()
    check(lib.vvae_philox_fill(ptr(out), out.numel(), int(seed), int(offset), 0 if kind == "normal" else 1, stream()),
          "vvae_philox_fill")
    return out


def sumsq_accum(g, out1, partials=None):
    """out1 += sum g^2.  With ``partials`` (fp32 scratch of sumsq_partials(n) elements) the summation order is fixed."""
    if partials is not None:
        check(lib.vvae_sumsq_f32_det(ptr(g), g.numel(), ptr(partials), ptr(out1), stream()), "vvae_sumsq_f32_det")
    else:
        check(lib.vvae_sumsq_f32(ptr(g), g.numel(), ptr(out1), stream()), "vvae_sumsq_f32")


def sumsq_partials(n):
    return int(lib.vvae_sumsq_partials(int(n)))


def adam_step_(p, g, m, v, lr, b1, b2, eps, step, gnorm_sq=None, clip=1.0, grad_scale=1.0, shadow=None):
    """``shadow``: optional bf16 tensor of p's size that receives the updated parameters (same rounding as vvae_cast)."""
    if shadow is not None:
        assert shadow.dtype == torch.bfloat16 and shadow.numel() == p.numel() and shadow.is_contiguous()
    check(lib.vvae_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(shadow), p.numel(), lr, b1, b2, eps, step, ptr(gnorm_sq),
                             clip, grad_scale, stream()), "vvae_adam_step")
