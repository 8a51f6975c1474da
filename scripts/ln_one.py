"""LayerNorm backward at [32768, 768] bf16, cold L2: one launch per variant (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
N, D = 32768, 768
x = torch.randn(N, D, device="cuda", generator=g).bfloat16()
dy = torch.randn(N, D, device="cuda", generator=g).bfloat16()
dres = torch.randn(N, D, device="cuda", generator=g).bfloat16()
gamma = torch.randn(D, device="cuda", generator=g); beta = torch.randn(D, device="cuda", generator=g)
dgam = torch.zeros(D, device="cuda"); dbet = torch.zeros(D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
for mode in (0, 2, 1):
    _ffi.lib.vvae_debug_set(7, mode * 17)
    flush.zero_()
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dgam, dbet)
    flush.zero_()
    ops.layernorm_fwd(x, gamma, beta)
torch.cuda.synchronize()
print("ok")
