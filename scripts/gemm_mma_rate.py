"""How fast does the tensor core retire 256x256x16 cta_group::2 MMAs (SS operands, SWIZZLE_128B) -- with and without the
rest of the kernel?  Uses the ablation bits of vvae_debug_set(10) and the in-kernel clock64 / globaltimer counters."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 256 * 74 * 24, 1536, 768
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for tB, what in ((False, "B MN-major (forward, W (in,out))"), (True, "B K-major (dgrad)")):
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g).bfloat16()
    for bits, name in ((16, "full kernel"), (16 | 1 | 2, "no TMA loads (barriers only)"), (16 | 4, "issuer never waits for operands"),
                       (16 | 4 | 8, "pure MMA stream (no operand waits, accumulators never drained)")):
        _ffi.lib.vvae_debug_set(10, bits)
        for _ in range(2):
            ops.gemm(A, B, transB=tB, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(A, B, transB=tB, out=out); e1.record(); torch.cuda.synchronize()
        buf = (C.c_ulonglong * 4)()
        _ffi.lib.vvae_debug_get(0, buf)
        ticks, ns, n = buf[0], buf[1], buf[2]
        print(json.dumps({"B": what, "mode": name, "kernel_us": round(e0.elapsed_time(e1) * 1e3, 1), "mma_issued_by_cta0": n,
                          "cycles_per_mma": round(ticks / max(n, 1), 1), "sm_clock_ghz": round(ticks / max(ns, 1), 3),
                          "tensor_util_at_clock": round(128.0 * n / max(ticks, 1), 3)}), flush=True)
    _ffi.lib.vvae_debug_set(10, 0)
