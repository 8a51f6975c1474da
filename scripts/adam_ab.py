"""vvae_adam_step on the production parameter count (170.5 M fp32): time of the fused clip + Adam + bf16 shadow pass."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
n = 170_518_528
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.randn(n, device="cuda", generator=g); gr = torch.randn(n, device="cuda", generator=g)
m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
sh = torch.empty(n, dtype=torch.bfloat16, device="cuda")
gsq = torch.ones(1, device="cuda")
def t(fn, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
fused = t(lambda: ops.adam_step_(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-8, 2, gsq, 1.0, 1.0, shadow=sh))
plain = t(lambda: ops.adam_step_(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-8, 2, gsq, 1.0, 1.0))
cast = t(lambda: ops.cast_into(p, sh))
scal = t(lambda: ops.adam_step_(p[1:], gr[1:], m[1:], v[1:], 1e-4, 0.9, 0.999, 1e-8, 2, gsq, 1.0, 1.0))
print(json.dumps({"n": n, "adam_vec4_with_shadow_ms": round(fused, 3), "GBps": round(n * 30 / fused / 1e6, 0),
                  "adam_vec4_ms": round(plain, 3), "cast_ms": round(cast, 3), "adam_scalar_ms": round(scal, 3)}))
