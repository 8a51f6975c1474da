"""clock64 timeline of one CTA of the persistent L = 256 attention forward (vvae_debug_set(10, 16), slots in the kernel's
comment): cycles relative to the issuer starting the CTA's 2nd unit."""
import ctypes as C, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
from video_vae_b200.ops import AttnGeom
_ffi.require_device()
H, HD = 8, 64
Q = H * HD
nseq, L = 128, 256
N = nseq * L
g = torch.Generator(device="cuda").manual_seed(0)
qk = (torch.randn(N, 2 * Q, device="cuda", generator=g) * 1.5).bfloat16()
v = torch.randn(N, 3 * Q, device="cuda", generator=g).bfloat16()[:, 2 * Q:]
geom = AttnGeom(nseq, 1, L, L, 0, 1)
sc = 1.0 / math.sqrt(HD)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
buf = (C.c_ulonglong * 32)()
names = ["issuer: unit start", "issuer: P.V_A and next S_A issued", "issuer: P_A seen", "issuer: P_B seen", "A: S seen", "A: row max", "A: P stored", "A: O seen", "A: O store issued",
         "B: S seen", "B: row max", "B: P stored", "B: O seen", "B: O store issued"]
for cta in (3, 77, 140):
    _ffi.lib.vvae_debug_set(10, 16)
    _ffi.lib.vvae_debug_set(0, cta)
    flush.zero_()
    ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], v, None, sc)
    torch.cuda.synchronize()
    _ffi.lib.vvae_debug_get(1, buf)
    t = list(buf)
    rec = {"cta": cta}
    for unit in (0, 1):
        rec["unit%d" % (unit + 1)] = {names[k]: int(t[14 * unit + k] - t[0]) for k in range(14) if t[14 * unit + k]}
    print(json.dumps(rec), flush=True)
_ffi.lib.vvae_debug_set(10, 0)
