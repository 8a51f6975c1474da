"""Bring-up harness for the tcgen05 GEMM (run on the B200 box: `python tests/gpu_bringup_gemm.py`).

Each case runs in its own subprocess so a faulting / hanging kernel cannot take the rest of the sweep down.
Results are appended to gpurun_out/gemm_bringup.jsonl.
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


class GemmArgs(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("A", C.c_void_p), ("lda", C.c_longlong), ("transA", C.c_int),
                ("B", C.c_void_p), ("ldb", C.c_longlong), ("transB", C.c_int),
                ("C", C.c_void_p), ("ldc", C.c_longlong),
                ("dtype", C.c_int), ("out_dtype", C.c_int),
                ("bias", C.c_void_p), ("epilogue", C.c_int),
                ("aux_in", C.c_void_p), ("ld_aux_in", C.c_longlong),
                ("aux_out", C.c_void_p), ("ld_aux_out", C.c_longlong),
                ("accumulate", C.c_int), ("backend", C.c_int), ("bsum_accum", C.c_void_p)]


def run_case(cfg):
    import torch
    lib = C.CDLL(os.path.join(ROOT, "video_vae_b200", "libvvae.so"))
    lib.vvae_last_error.restype = C.c_char_p
    lib.vvae_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    lib.vvae_debug_set.argtypes = [C.c_int, C.c_longlong]
    for k, v in cfg.get("dbg", {}).items():
        lib.vvae_debug_set(int(k), int(v))
    M, N, K = cfg["M"], cfg["N"], cfg["K"]
    tA, tB = cfg["tA"], cfg["tB"]
    mode = cfg.get("mode", 0)
    out_f32 = cfg.get("out_f32", 0)
    acc = cfg.get("acc", 0)
    backend = cfg.get("backend", 2)
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn((K, M) if tA else (M, K), generator=g, device=dev) * 0.5).bfloat16()
    B = (torch.randn((N, K) if tB else (K, N), generator=g, device=dev) * 0.5).bfloat16()
    bias = torch.randn(N, generator=g, device=dev) if cfg.get("bias", 0) else None
    aux_in = (torch.randn(M, N, generator=g, device=dev)).bfloat16() if mode in (2, 3) else None
    aux_out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16) if mode == 1 else None
    Cm = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    if acc:
        Cm += 1.0
    bsum = torch.zeros(N, device=dev) if cfg.get("bsum", 0) else None
    a = GemmArgs(M, N, K, A.data_ptr(), A.stride(0), tA, B.data_ptr(), B.stride(0), tB, Cm.data_ptr(), N,
                 1, 0 if out_f32 else 1, bias.data_ptr() if bias is not None else None, mode,
                 aux_in.data_ptr() if aux_in is not None else None, N,
                 aux_out.data_ptr() if aux_out is not None else None, N, acc, backend,
                 bsum.data_ptr() if bsum is not None else None)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vvae_gemm(C.byref(a), st)
    if rc != 0:
        return {"rc": rc, "err": lib.vvae_last_error().decode()}
    torch.cuda.synchronize()
    opA = A.float().t() if tA else A.float()
    opB = B.float().t() if tB else B.float()
    ref = opA @ opB
    if bias is not None:
        ref = ref + bias
    res = {"rc": 0}
    if bsum is not None:
        refb = B.float().sum(0)
        res["bsum_err"] = ((bsum - refb).abs().max() / refb.abs().max()).item()
    if mode == 1:
        res["aux_err"] = (aux_out.float() - ref).abs().max().item()
        ref = torch.nn.functional.silu(ref.bfloat16().float())
    elif mode == 2:
        ref = ref + aux_in.float()
    elif mode == 3:
        x = aux_in.float()
        s = torch.sigmoid(x)
        ref = ref * (s * (1 + x * (1 - s)))
    if acc:
        ref = ref + 1.0
    if cfg.get("row_shift"):
        rs = cfg["row_shift"]
        # rows of each 128-row tile should equal the reference rows shifted down by rs (valid while inside the tile)
        keep = 128 - rs
        err_s = (Cm.float()[:keep] - ref[rs:rs + keep]).abs()
        res["shift_max_abs_err"] = err_s.max().item()
        res["shift_frac_bad"] = (err_s > 0.02 * ref[rs:rs + keep].abs() + 0.05).float().mean().item()
    err = (Cm.float() - ref).abs()
    res["max_abs_err"] = err.max().item()
    res["ref_absmax"] = ref.abs().max().item()
    res["frac_bad"] = (err > 0.02 * ref.abs() + 0.05).float().mean().item()
    if cfg.get("time", 0):
        if acc:
            pass
        for _ in range(3):
            lib.vvae_gemm(C.byref(a), st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            lib.vvae_gemm(C.byref(a), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        # cuBLAS (torch.matmul) comparator on the same operands
        e0.record()
        for _ in range(iters):
            torch.matmul(A.t() if tA else A, B.t() if tB else B)
        e1.record()
        torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / iters) / 1e9
    return res


CASES = []


def add(name, **kw):
    kw["name"] = name
    CASES.append(kw)


# correctness, small, every operand-major combination
add("KK_small", M=256, N=256, K=256, tA=0, tB=1)
add("KMN_small", M=256, N=256, K=256, tA=0, tB=0)
add("MNMN_small", M=256, N=256, K=256, tA=1, tB=0, out_f32=1)
add("MNK_small", M=256, N=256, K=256, tA=1, tB=1)
add("simt_ref", M=256, N=256, K=256, tA=0, tB=0, backend=1)
# alternative MN-major encodings in case the derivation is off (dbg: 1=lbo 2=sbo 3=kadv)
add("KMN_alt_swap", M=256, N=256, K=256, tA=0, tB=0, dbg={1: 1024, 2: 8192})
add("KMN_alt_kadv", M=256, N=256, K=256, tA=0, tB=0, dbg={3: 4096})
add("KMN_bn64", M=256, N=256, K=256, tA=0, tB=0, dbg={6: 64})
add("KMN_bn128", M=256, N=256, K=256, tA=0, tB=0, dbg={6: 128})
add("KK_bn64", M=256, N=256, K=256, tA=0, tB=1, dbg={6: 64})
# epilogues
add("fwd_bias", M=384, N=512, K=320, tA=0, tB=0, bias=1)
add("fwd_silu", M=384, N=512, K=320, tA=0, tB=0, bias=1, mode=1)
add("fwd_resid", M=384, N=512, K=320, tA=0, tB=0, bias=1, mode=2)
add("dgrad_dsilu", M=384, N=512, K=320, tA=0, tB=1, mode=3)
add("wgrad_acc", M=768, N=1536, K=4096, tA=1, tB=0, out_f32=1, acc=1)
# the same through single-CTA tiles (dbg 8 = 1 disables CTA pairs)
add("fwd_silu_cg1", M=384, N=512, K=320, tA=0, tB=0, bias=1, mode=1, dbg={8: 1})
add("fwd_resid_cg1", M=384, N=512, K=320, tA=0, tB=0, bias=1, mode=2, dbg={8: 1})
add("wgrad_acc_cg1", M=768, N=1536, K=4096, tA=1, tB=0, out_f32=1, acc=1, dbg={8: 1})
add("fwd_bias_n96", M=512, N=96, K=768, tA=0, tB=0, bias=1)
add("fwd_resid_big", M=4096, N=768, K=512, tA=0, tB=0, bias=1, mode=2)
add("fwd_silu_big", M=4096, N=1536, K=768, tA=0, tB=0, bias=1, mode=1)
# ragged
add("tail_M300_N192_K96", M=300, N=192, K=96, tA=0, tB=0, bias=1)
add("tail_dgrad", M=300, N=96 + 8, K=200, tA=0, tB=1)
add("tail_wgrad", M=200, N=136, K=1000, tA=1, tB=0, out_f32=1, acc=1)
# production shapes, timed
add("prod_qkv_fwd", M=32768, N=1536, K=768, tA=0, tB=0, bias=1, time=1)
add("prod_mlp_dn_fwd", M=32768, N=768, K=1536, tA=0, tB=0, bias=1, mode=2, time=1)
add("prod_dgrad", M=32768, N=768, K=1536, tA=0, tB=1, time=1)
add("prod_wgrad", M=768, N=1536, K=32768, tA=1, tB=0, out_f32=1, acc=1, time=1)
add("prod_out_fwd", M=32768, N=768, K=512, tA=0, tB=0, bias=1, mode=2, time=1)
add("prod_mlp_up_fwd", M=32768, N=1536, K=768, tA=0, tB=0, bias=1, mode=1, time=1)
add("prod_do_dgrad", M=32768, N=512, K=768, tA=0, tB=1, time=1)
add("prod_du_dgrad", M=32768, N=1536, K=768, tA=0, tB=1, mode=3, time=1)
add("prod_wgrad_o", M=512, N=768, K=32768, tA=1, tB=0, out_f32=1, acc=1, time=1)
add("prod_wgrad_w2", M=1536, N=768, K=32768, tA=1, tB=0, out_f32=1, acc=1, time=1)
add("wgrad_bsum_small", M=768, N=1536, K=4096, tA=1, tB=0, out_f32=1, acc=1, bsum=1)
add("wgrad_bsum_tail", M=200, N=136, K=1000, tA=1, tB=0, out_f32=1, acc=1, bsum=1)
add("wgrad_bsum_n96", M=768, N=96, K=2048, tA=1, tB=0, out_f32=1, acc=1, bsum=1)
add("prod_wgrad_bsum", M=768, N=1536, K=32768, tA=1, tB=0, out_f32=1, acc=1, time=1, bsum=1)
add("prod_wgrad_o_bsum", M=512, N=768, K=32768, tA=1, tB=0, out_f32=1, acc=1, time=1, bsum=1)
add("prod_qkv_fwd_cg1", M=32768, N=1536, K=768, tA=0, tB=0, bias=1, time=1, dbg={8: 1})
add("prod_mlp_up_fwd_cg1", M=32768, N=1536, K=768, tA=0, tB=0, bias=1, mode=1, time=1, dbg={8: 1})
add("prod_qkv_fwd_bn128", M=32768, N=1536, K=768, tA=0, tB=0, bias=1, time=1, dbg={6: 128})


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        cfg = json.loads(sys.argv[2])
        print("RESULT " + json.dumps(run_case(cfg)))
        return
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    log = open(os.path.join(OUT, "gemm_bringup.jsonl"), "a")
    for cfg in CASES:
        if only and cfg["name"] not in only:
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", json.dumps(cfg)], capture_output=True, text=True,
                               timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            res = json.loads(line[-1][7:]) if line else {"rc": -1, "stderr": r.stderr[-600:], "code": r.returncode}
        except subprocess.TimeoutExpired:
            res = {"rc": -2, "err": "timeout (hang)"}
        res["name"] = cfg["name"]
        res["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(res), flush=True)
        log.write(json.dumps(res) + "\n")
        log.flush()


if __name__ == "__main__":
    main()
