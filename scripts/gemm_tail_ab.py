#!/usr/bin/env python
"""Partial-last-wave column slicing of the tcgen05 GEMM (gemm_sm100.cu, Sm100Params::tail_s) against whole tiles
(vvae_debug_set(17, 1)) on the production shapes: results must be bit-identical, times are cold-L2 medians."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402
from scripts.gemm_probe import CASES  # noqa: E402

_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
EXTRA = [("odd M: 8 x 13 x 256 tokens", 26624, 768, 1536, False, True, _ffi.EPI_NONE, False, False, False),
         ("odd M residual", 26624 + 128, 768, 512, False, False, _ffi.EPI_RESIDUAL, True, False, False)]
tot = {1: 0.0, 0: 0.0}
for name, M, N, K, tA, tB, epi, bias, acc, bsum in CASES + EXTRA:
    if acc:
        continue
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g).bfloat16()
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    aux_in = torch.randn(M, N, device="cuda", generator=g).bfloat16() if epi in (_ffi.EPI_RESIDUAL, _ffi.EPI_DSILU) else None
    res = {}
    for whole in (1, 0, 2):
        _ffi.lib.vvae_debug_set(17, whole)
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        aux_out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if epi == _ffi.EPI_SILU else None

        def run():
            ops.gemm(A, B, transA=tA, transB=tB, out=out, bias=b, epilogue=epi, aux_in=aux_in, aux_out=aux_out)
        run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[whole] = (out.clone(), None if aux_out is None else aux_out.clone(), sorted(ts)[3])
    same = torch.equal(res[0][0], res[1][0]) and (res[0][1] is None or torch.equal(res[0][1], res[1][1]))
    if name in [c[0] for c in CASES]:
        tot[1] += res[1][2]; tot[0] += res[0][2]
    print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "whole_tiles_us": round(res[1][2], 1),
                      "sliced_tail_us": round(res[0][2], 1), "two_slices_us": round(res[2][2], 1), "bit_identical": same}), flush=True)
_ffi.lib.vvae_debug_set(17, 0)
print(json.dumps({"sum_whole_us": round(tot[1], 1), "sum_sliced_us": round(tot[0], 1)}))
