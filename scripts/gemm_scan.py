"""GEMM time vs number of tile waves (fixed N, K; M varies): separates the per-launch fixed cost from the per-wave cost."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
N, K = 1536, 768
B = torch.randn(K, N, device="cuda", generator=g).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
for M in (256 * 74, 256 * 74 * 2, 256 * 74 * 3, 256 * 74 * 4, 256 * 74 * 6, 256 * 74 * 8, 256 * 74 * 12, 32768):
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, out=out, bias=b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(A, B, out=out, bias=b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = ((M + 255) // 256) * (N // 256)
    print(json.dumps({"M": M, "tiles": tiles, "waves": round(tiles / 74, 2), "us": round(ms * 1e3, 2),
                      "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}), flush=True)
