#!/usr/bin/env python
"""GroupNorm+SiLU backward on the U-Net's map shapes, cold L2: 8 channels per thread (vvae_debug_set(7, 0x400)) against
4 channels per thread for the narrow maps; compares the results of the two."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


for (H, C) in ((256, 16), (128, 32), (64, 64)):
    x = torch.randn(8, 16, H, H, C, device="cuda", generator=g).bfloat16()
    dy = torch.randn(8, 16, H, H, C, device="cuda", generator=g).bfloat16()
    ga = torch.randn(C, device="cuda", generator=g); be = torch.randn(C, device="cuda", generator=g)
    y, mean, rstd = ops.groupnorm_silu_fwd(x, ga, be, 8)
    out = {"map": [8, 16, H, H, C], "bytes_bwd": x.numel() * 2 * 5}
    res = {}
    for mode, nm in ((0x400, "v8"), (0, "v4")):
        _ffi.lib.vvae_debug_set(7, mode)
        dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda"); cs = torch.zeros(C, device="cuda")
        dx = ops.groupnorm_silu_bwd(dy, C, x, ga, be, mean, rstd, dg, db, 8, dx_colsum=cs)
        torch.cuda.synchronize()
        res[nm] = (dx.float(), dg.clone(), db.clone(), cs.clone())
        out[nm + "_us"] = round(timed(lambda: ops.groupnorm_silu_bwd(dy, C, x, ga, be, mean, rstd, dg, db, 8, dx_colsum=cs)), 1)
    _ffi.lib.vvae_debug_set(7, 0)
    for i, nm in enumerate(("dx", "dgamma", "dbeta", "dx_colsum")):
        a, b = res["v4"][i], res["v8"][i]
        out[nm + "_maxrel"] = float(((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item())
    yres = {}
    for mode, nm in ((0x400, "v8"), (0, "v4")):
        _ffi.lib.vvae_debug_set(7, mode)
        yy, mm, rr = ops.groupnorm_silu_fwd(x, ga, be, 8)
        yres[nm] = (yy.float(), mm.clone(), rr.clone())
        out["fwd_" + nm + "_us"] = round(timed(lambda: ops.groupnorm_silu_fwd(x, ga, be, 8)), 1)
    _ffi.lib.vvae_debug_set(7, 0)
    out["fwd_y_maxrel"] = float(((yres["v4"][0] - yres["v8"][0]).abs().max() / yres["v8"][0].abs().max()).item())
    out["fwd_mean_maxrel"] = float(((yres["v4"][1] - yres["v8"][1]).abs().max() / yres["v8"][1].abs().max()).item())
    print(json.dumps(out), flush=True)
