"""One launch each of the HBM-bound normalisation kernels at BASELINE configs[1] shapes (for ncu): LayerNorm fwd/bwd over
[32768, 768], QK-norm+RoPE fwd/bwd over [32768, 1536], GroupNorm+SiLU fwd/bwd over [8,16,256,256,16]."""
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
for it in range(2):
    flush.zero_()
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    flush.zero_()
    dx = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dgam, dbet)
H, HD = 8, 64
qkv = torch.randn(N, 3 * H * HD, device="cuda", generator=g).bfloat16()
dqkv = torch.randn(N, 3 * H * HD, device="cuda", generator=g).bfloat16()
qs = torch.randn(HD, device="cuda", generator=g); ks = torch.randn(HD, device="cuda", generator=g)
pos = torch.arange(256, device="cuda").float()[:, None] * torch.exp(-torch.arange(0, HD, 2, device="cuda").float() / HD * 9.2)[None]
emb = torch.cat([pos, pos], -1)
cos, sin = torch.cos(emb).bfloat16().contiguous(), torch.sin(emb).bfloat16().contiguous()
dqs = torch.zeros(HD, device="cuda"); dks = torch.zeros(HD, device="cuda"); dbqk = torch.zeros(2 * H * HD, device="cuda")
for it in range(2):
    flush.zero_()
    qk = ops.qknorm_rope_fwd(qkv, qs, ks, cos, sin, H, HD, 1, 256)
    flush.zero_()
    ops.qknorm_rope_bwd_(dqkv, qkv, qs, ks, cos, sin, dqs, dks, H, HD, 1, 256, dbias_qk=dbqk)
xg = torch.randn(8, 16, 256, 256, 16, device="cuda", generator=g).bfloat16()
dyg = torch.randn(8, 16, 256, 256, 16, device="cuda", generator=g).bfloat16()
gg = torch.randn(16, device="cuda", generator=g); bg = torch.randn(16, device="cuda", generator=g)
dgg = torch.zeros(16, device="cuda"); dbg = torch.zeros(16, device="cuda"); cs = torch.zeros(16, device="cuda")
for it in range(2):
    flush.zero_()
    yg, mg, rg = ops.groupnorm_silu_fwd(xg, gg, bg, 8)
    flush.zero_()
    dxg = ops.groupnorm_silu_bwd(dyg, 16, xg, gg, bg, mg, rg, dgg, dbg, 8, dx_colsum=cs)
torch.cuda.synchronize()
print("ok")
