"""Bring-up: tensor-core ConvTranspose (1,2,2) (GEMM + 5-D TMA views) vs the generic kernel.  `small` = tiny shapes only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import ops  # noqa: E402
from video_vae_b200._ffi import lib, check, ptr, dt, stream  # noqa: E402

SMALL = len(sys.argv) > 1 and sys.argv[1] == "small"
ONLY = sys.argv[2] if len(sys.argv) > 2 else None
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9)).item()


def run(name, B, T, H, W, Cin, Cout, timed=False, parts="fdw"):
    x = torch.randn(B, T, H, W, Cin, device=dev, generator=g).bfloat16()
    w = (torch.randn(1, 2, 2, Cin, Cout, device=dev, generator=g) * 0.1).bfloat16()
    bias = torch.randn(Cout, device=dev, generator=g)
    ld = 2 * Cout
    res = {"name": name}
    outs = {}
    for mode in ("generic", "tc"):
        cat = torch.zeros(B, T, 2 * H, 2 * W, ld, device=dev, dtype=torch.bfloat16)
        ws = ops._convt_workspace(B * T, H, W, Cin, Cout, x.device) if mode == "tc" else None
        wsn = ws.numel() if ws is not None else 0
        if "f" in parts:
            check(lib.vvae_convT122_fwd(ptr(x), ptr(w), ptr(bias), ptr(cat), ld, B * T, H, W, Cin, Cout, dt(x), ptr(ws), wsn,
                                        stream()), "fwd")
        dcat = torch.randn(B, T, 2 * H, 2 * W, ld, device=dev, generator=torch.Generator(device=dev).manual_seed(5)).bfloat16()
        dx = torch.zeros_like(x)
        dw = torch.zeros(1, 2, 2, Cin, Cout, device=dev)
        if "d" in parts or "w" in parts:
            check(lib.vvae_convT122_bwd(ptr(dcat), ld, ptr(x), ptr(w), ptr(dx) if "d" in parts else None,
                                        ptr(dw) if "w" in parts else None, B * T, H, W, Cin, Cout, dt(x), ptr(ws), wsn,
                                        stream()), "bwd")
        torch.cuda.synchronize()
        outs[mode] = (cat[..., :Cout].float().clone(), dx.float().clone(), dw.clone())
        if timed and mode == "tc":
            for kind in ("fwd", "bwd"):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    if kind == "fwd":
                        lib.vvae_convT122_fwd(ptr(x), ptr(w), ptr(bias), ptr(cat), ld, B * T, H, W, Cin, Cout, dt(x), ptr(ws),
                                              wsn, stream())
                    else:
                        lib.vvae_convT122_bwd(ptr(dcat), ld, ptr(x), ptr(w), ptr(dx), ptr(dw), B * T, H, W, Cin, Cout, dt(x),
                                              ptr(ws), wsn, stream())
                e1.record()
                torch.cuda.synchronize()
                res[kind + "_ms"] = e0.elapsed_time(e1) / 5
    res["y_err"] = rel(outs["tc"][0], outs["generic"][0])
    res["dx_err"] = rel(outs["tc"][1], outs["generic"][1]) if "d" in parts else None
    res["dw_err"] = rel(outs["tc"][2], outs["generic"][2]) if "w" in parts else None
    print(json.dumps(res), flush=True)


parts = ONLY or "fdw"
run("t8_32_16", 2, 2, 8, 8, 32, 16, parts=parts)
run("t16_64_32", 1, 2, 8, 16, 64, 32, parts=parts)
run("t8_128_64", 2, 2, 8, 8, 128, 64, parts=parts)
if not SMALL:
    run("w128_32_16_tail", 1, 3, 5, 128, 32, 16, parts=parts)
    run("prod_dec2", 8, 16, 128, 128, 32, 16, timed=True)
    run("prod_dec1", 8, 16, 64, 64, 64, 32, timed=True)
    run("prod_dec0", 8, 16, 32, 32, 128, 64, timed=True)
