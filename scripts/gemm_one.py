import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 32768, 1536, 768
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
B = torch.randn(K, N, device="cuda", generator=g).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(A, B, out=out, bias=b)
torch.cuda.synchronize()
print("ok")
