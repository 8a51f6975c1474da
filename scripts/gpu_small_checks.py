"""Spot checks: (1) weight gradient with M = 96 on the tcgen05 path against the SIMT kernel; (2) pitched pixel shuffle
round trip; (3) UnembedFn + UNet in bf16 with the pitched feature map against the packed path (same numbers)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
K, M, N = 32768, 96, 768
X = torch.randn(K, M, device="cuda", generator=g).bfloat16()
dY = torch.randn(K, N, device="cuda", generator=g).bfloat16()
outs = {}
for be in (_ffi.BACKEND_SIMT, _ffi.BACKEND_AUTO):
    dw = torch.zeros(M, N, device="cuda"); bs = torch.zeros(N, device="cuda")
    ops.gemm(X, dY, transA=True, out=dw, accumulate=True, bsum=bs, backend=be)
    torch.cuda.synchronize()
    outs[be] = (dw, bs)
a, b = outs[_ffi.BACKEND_AUTO], outs[_ffi.BACKEND_SIMT]
print(json.dumps({"wgrad_M96_rel": ((a[0] - b[0]).abs().max() / b[0].abs().max()).item(),
                  "bsum_rel": ((a[1] - b[1]).abs().max() / b[1].abs().max()).item()}))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dw = torch.zeros(M, N, device="cuda"); bs = torch.zeros(N, device="cuda")
e0.record()
for _ in range(5):
    ops.gemm(X, dY, transA=True, out=dw, accumulate=True, bsum=bs)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"wgrad_M96_us": e0.elapsed_time(e1) * 200}))
# (2)
bt, H, W, CU, P = 6, 64, 64, 12, 16
tok = torch.randn(bt * (H // P) * (W // P), P * P * CU, device="cuda", generator=g).bfloat16()
packed = ops.pixel_shuffle(tok, bt, H, W, CU, P, to_tokens=False)
pitched = ops.pixel_shuffle(tok, bt, H, W, CU, P, to_tokens=False, vox_ld=16)
back = ops.pixel_shuffle(pitched, bt, H, W, CU, P, to_tokens=True, vox_ld=16)
print(json.dumps({"pitched_equals_packed": bool(torch.equal(pitched[..., :CU], packed)),
                  "pads_zero": bool((pitched[..., CU:] == 0).all()), "round_trip": bool(torch.equal(back, tok))}))
