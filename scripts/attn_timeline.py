"""clock64 timeline of single CTAs of the tcgen05 attention backward at the production spatial shape (128 sequences of
L = 256, 8 heads x 64): vvae_debug_set(10, 16) + vvae_debug_get(1, .).  One JSON line per probed CTA, cycles relative
to the CTA's start.  Slots: 0 start, 1 TMEM allocated, 2 after griddepcontrol.wait, 3 MMA warp: block 0 landed,
4-7 MMA warp: P/dS of tile t ready, 8-11 MMA warp: output contractions of tile t issued, 12 softmax warp: D = rowsum(dO o O)
done, 13-16 softmax warp: S/dP of tile t ready, 17-20 softmax warp: P/dS of tile t computed (registers),
21-24 softmax warp: P/dS of tile t stored + arrived, 25-26 dK/dV of key block j drained, 27 CTA end, 31 SM id."""
import ctypes as C
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402
from video_vae_b200.ops import AttnGeom  # noqa: E402

H, HD = 8, 64
Q = H * HD
_ffi.require_device()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
b, t, hw = 8, 16, 256
N = b * t * hw
qkv = torch.randn(N, 3 * Q, device=dev, generator=g).bfloat16()
qk = (torch.randn(N, 2 * Q, device=dev, generator=g) * 1.5).bfloat16()
d_o = torch.randn(N, Q, device=dev, generator=g).bfloat16()
geom = AttnGeom(b * t, 1, hw, hw, 0, 1)
sc = 1.0 / math.sqrt(HD)
o, lse = ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], None, sc)
dqkv = torch.zeros(N, 3 * Q, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bwd():
    ops.attn_bwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], o, lse, d_o, dqkv[:, :Q], dqkv[:, Q:2 * Q],
                 dqkv[:, 2 * Q:], None, sc)


def timed(reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bwd(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


res = {}
for old in (1, 0):
    _ffi.lib.vvae_debug_set(16, old)
    dqkv.zero_()
    bwd()
    torch.cuda.synchronize()
    res[old] = dqkv.float().clone()
    print(json.dumps({"kernel": "one unit per CTA" if old else "persistent", "bwd_us_cold_L2": round(timed(), 1)}), flush=True)
for nm, sl in (("dq", slice(0, Q)), ("dk", slice(Q, 2 * Q)), ("dv", slice(2 * Q, 3 * Q))):
    a_, b_ = res[0][:, sl], res[1][:, sl]
    print(json.dumps({"persistent_vs_one_unit": nm, "max_rel": ((a_ - b_).abs().max() / b_.abs().max()).item()}), flush=True)
# persistent kernel (slots, relative to the issuer's start of the CTA's 2nd unit): 0-3 issuer starts tile t, 4-7 issuer sees
# P/dS of tile t, 8-11 contractions of tile t issued, 12-15 softmax warp sees S/dP of tile t, 16-19 P/dS computed,
# 20-23 P/dS stored, 24 dK0/dV0 drained, 25 dQ0 drained, 26 dK1/dV1/dQ1 drained
buf = (C.c_ulonglong * 32)()
for abl in (1, 2):
    _ffi.lib.vvae_debug_set(10, abl)
    print(json.dumps({"ablation": {1: "drain warps: TMEM loads only", 2: "no L2 prefetch of the next unit"}[abl],
                      "bwd_us_cold_L2": round(timed(), 1)}), flush=True)
for cta, abl in ((5, 0), (100, 0), (5, 2)):
    _ffi.lib.vvae_debug_set(10, 16 + abl)
    _ffi.lib.vvae_debug_set(0, cta)
    flush.zero_()
    bwd()
    torch.cuda.synchronize()
    _ffi.lib.vvae_debug_get(1, buf)
    v = list(buf)
    print(json.dumps({"persistent_cta": cta, "ablation": abl, "cycles": {i: int(v[i] - v[0]) for i in range(1, 27) if v[i]}}), flush=True)
_ffi.lib.vvae_debug_set(16, 1)
for cta in (700,):
    _ffi.lib.vvae_debug_set(0, cta)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bwd(); e1.record()
    torch.cuda.synchronize()
    _ffi.lib.vvae_debug_get(1, buf)
    v = list(buf)
    rel = {i: int(v[i] - v[0]) for i in range(1, 28) if v[i]}
    print(json.dumps({"cta": cta, "sm": int(v[31]), "kernel_us": round(e0.elapsed_time(e1) * 1e3, 1), "cycles": rel}), flush=True)
_ffi.lib.vvae_debug_set(10, 0)
_ffi.lib.vvae_debug_set(0, 0)
_ffi.lib.vvae_debug_set(16, 0)

# ---- forward kernel (slots: 0 start, 1 after griddepcontrol.wait, 2 Q|K landed (issuer), 3 P ready (issuer), 4 PV issued,
# 5 softmax warp: S ready, 6 row max done, 7 P written, 8 O ready, 9 O stored)
_ffi.lib.vvae_debug_set(10, 16)
for cta in (0, 300, 1000, 2000):
    _ffi.lib.vvae_debug_set(0, cta)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], None, sc); e1.record()
    torch.cuda.synchronize()
    _ffi.lib.vvae_debug_get(1, buf)
    v = list(buf)
    print(json.dumps({"fwd_cta": cta, "sm": int(v[31]), "kernel_us": round(e0.elapsed_time(e1) * 1e3, 1),
                      "cycles": {i: int(v[i] - v[0]) for i in range(1, 10) if v[i]}}), flush=True)
_ffi.lib.vvae_debug_set(10, 0)
_ffi.lib.vvae_debug_set(0, 0)
