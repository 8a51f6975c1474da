#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM at BASELINE configs[1] shapes (N = 32768 tokens).

    python scripts/gemm_probe.py   -> one JSON line per (shape, mode): ms, TFLOP/s, fraction of the measured bf16 peaks
Each GEMM is timed over 5 back-to-back launches after one warm-up (CUDA events); operands are 25-200 MB each.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402
from video_vae_b200._ffi import EPI_DSILU, EPI_NONE, EPI_RESIDUAL, EPI_SILU  # noqa: E402

NTOK = 32768
CASES = [
    # name, M, N, K, transA, transB, epilogue, bias, accumulate(fp32 out), bsum
    ("qkv_fwd        X[N,768]  W[768,1536] +b", NTOK, 1536, 768, False, False, EPI_NONE, True, False, False),
    ("mlp_up_fwd     +b SiLU (2 outputs)", NTOK, 1536, 768, False, False, EPI_SILU, True, False, False),
    ("out_proj_fwd   X[N,512]  W[512,768] +b +res", NTOK, 768, 512, False, False, EPI_RESIDUAL, True, False, False),
    ("mlp_down_fwd   X[N,1536] W[1536,768] +b +res", NTOK, 768, 1536, False, False, EPI_RESIDUAL, True, False, False),
    ("qkv_dgrad      dY[N,1536] W^T -> [N,768]", NTOK, 768, 1536, False, True, EPI_NONE, False, False, False),
    ("mlp_up_dgrad   dU[N,1536] W^T -> [N,768]", NTOK, 768, 1536, False, True, EPI_NONE, False, False, False),
    ("out_proj_dgrad dY[N,768] W^T -> [N,512]", NTOK, 512, 768, False, True, EPI_NONE, False, False, False),
    ("mlp_down_dgrad dY[N,768] W^T -> [N,1536] * dSiLU", NTOK, 1536, 768, False, True, EPI_DSILU, False, False, False),
    ("qkv_wgrad      X^T[768,N] dY[N,1536] (+bias grad)", 768, 1536, NTOK, True, False, EPI_NONE, False, True, True),
    ("mlp_down_wgrad A^T[1536,N] dY[N,768] (+bias grad)", 1536, 768, NTOK, True, False, EPI_NONE, False, True, True),
    ("out_proj_wgrad O^T[512,N] dY[N,768] (+bias grad)", 512, 768, NTOK, True, False, EPI_NONE, False, True, True),
    ("unembed_up_fwd X[N,768] W[768,3072] +b", NTOK, 3072, 768, False, False, EPI_NONE, True, False, False),
]


def main():
    _ffi.require_device()
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    g = torch.Generator(device="cuda").manual_seed(0)
    tot = 0.0
    for name, M, N, K, tA, tB, epi, bias, acc, bsum in CASES:
        A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g).bfloat16()
        B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g).bfloat16()
        b = torch.randn(N, device="cuda", generator=g) if bias else None
        aux_in = torch.randn(M, N, device="cuda", generator=g).bfloat16() if epi in (EPI_RESIDUAL, EPI_DSILU) else None
        aux_out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if epi == EPI_SILU else None
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if acc else torch.bfloat16)
        bs = torch.zeros(N, device="cuda") if bsum else None

        def run():
            ops.gemm(A, B, transA=tA, transB=tB, out=out, bias=b, epilogue=epi, aux_in=aux_in, aux_out=aux_out,
                     accumulate=acc, bsum=bs)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        tf = 2.0 * M * N * K / ms / 1e9
        tot += ms
        print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "ms": round(ms, 4), "tflops": round(tf, 1),
                          "frac_of_burst_peak": round(tf / peaks["bf16_tflops"], 3),
                          "frac_of_sustained_peak": round(tf / peaks["bf16_tflops_sustained"], 3)}), flush=True)
    print(json.dumps({"gemm": "sum of the 12 cases", "ms": round(tot, 3)}))


if __name__ == "__main__":
    main()
