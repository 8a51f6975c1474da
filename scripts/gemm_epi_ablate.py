#!/usr/bin/env python
"""Upper bound of any epilogue improvement: every production GEMM shape with vvae_debug_set(10, 8) (accumulators never
drained, nothing stored: TIMING only) against the full kernel; cold L2."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402
from scripts.gemm_probe import CASES  # noqa: E402
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = {0: 0.0, 8: 0.0}
for name, M, N, K, tA, tB, epi, bias, acc, bsum in CASES:
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g).bfloat16()
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    aux_in = torch.randn(M, N, device="cuda", generator=g).bfloat16() if epi in (_ffi.EPI_RESIDUAL, _ffi.EPI_DSILU) else None
    aux_out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if epi == _ffi.EPI_SILU else None
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if acc else torch.bfloat16)
    bs = torch.zeros(N, device="cuda") if bsum else None
    res = {}
    for bits in (0, 8):
        _ffi.lib.vvae_debug_set(10, bits)

        def run():
            ops.gemm(A, B, transA=tA, transB=tB, out=out, bias=b, epilogue=epi, aux_in=aux_in, aux_out=aux_out, accumulate=acc, bsum=bs)
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[bits] = sorted(ts)[3]
        tot[bits] += res[bits]
    _ffi.lib.vvae_debug_set(10, 0)
    print(json.dumps({"gemm": name, "full_us": round(res[0], 1), "no_epilogue_us": round(res[8], 1)}), flush=True)
print(json.dumps({"sum_full_us": round(tot[0], 1), "sum_no_epilogue_us": round(tot[8], 1)}))
