"""Probe: 192-column GEMM tiles for the N = 768 dgrads (debug key 12) against the default 256-column tiles."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 32768, 768, 1536
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()      # dgrad: dX = dY . W^T with W [in=N, out=K]
A2 = A.clone()
ref = None
for key in (0, 1, 0, 1):
    _ffi.lib.vvae_debug_set(12, key)
    out = ops.gemm(A, W, transB=True)
    torch.cuda.synchronize()
    if ref is None:
        ref = out.float()
        exact = (A[:4096].float() @ W.float().t())
        base_err = ((ref[:4096] - exact).abs().max() / exact.abs().max()).item()
    err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
    for _ in range(3):
        ops.gemm(A, W, transB=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        ops.gemm(A if i % 2 == 0 else A2, W, transB=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"bn": 192 if key else 256, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "err_vs_bn256": err,
                      "bn256_err_vs_fp32": base_err}), flush=True)
_ffi.lib.vvae_debug_set(12, 0)
