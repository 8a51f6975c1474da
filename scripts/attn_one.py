"""One launch each of the attention kernels at the production spatial shape (128 sequences x 8 heads x L = 256), cold L2
(for ncu): forward, backward (persistent kernel + D pre-pass), and the one-unit-per-CTA backward for comparison."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
from video_vae_b200.ops import AttnGeom
_ffi.require_device()
H, HD = 8, 64
Q = H * HD
g = torch.Generator(device="cuda").manual_seed(0)
b, t, hw = 8, 16, 256
N = b * t * hw
qkv = torch.randn(N, 3 * Q, device="cuda", generator=g).bfloat16()
qk = (torch.randn(N, 2 * Q, device="cuda", generator=g) * 1.5).bfloat16()
d_o = torch.randn(N, Q, device="cuda", generator=g).bfloat16()
geom = AttnGeom(b * t, 1, hw, hw, 0, 1)
sc = 1.0 / math.sqrt(HD)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dqkv = torch.zeros(N, 3 * Q, device="cuda", dtype=torch.bfloat16)
for it in range(2):
    flush.zero_()
    o, lse = ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], None, sc)
    for old in (0, 1):
        _ffi.lib.vvae_debug_set(16, old)
        flush.zero_()
        ops.attn_bwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], o, lse, d_o, dqkv[:, :Q], dqkv[:, Q:2 * Q],
                     dqkv[:, 2 * Q:], None, sc)
_ffi.lib.vvae_debug_set(16, 0)
torch.cuda.synchronize()
print("ok")
