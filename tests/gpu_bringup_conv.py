"""Bring-up: tensor-core conv3d (fwd, dgrad) vs the generic kernel on the same bf16 data. Run on the B200 box."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402

CASES = [
    # name, B, T, H, W, Cin, Cout, ks, x_ld, y_ld
    ("c16_16_w256", 1, 3, 6, 256, 16, 16, (3, 3, 3), 16, 16),
    ("c12_16_w256_ld16", 1, 3, 6, 256, 12, 16, (3, 3, 3), 16, 16),
    ("c32_16_w256", 1, 3, 4, 256, 32, 16, (3, 3, 3), 32, 16),
    ("c16_32_w128", 1, 3, 6, 128, 16, 32, (3, 3, 3), 16, 32),
    ("c64_32_w128", 1, 2, 5, 128, 64, 32, (3, 3, 3), 64, 32),
    ("c32_64_w64", 1, 3, 8, 64, 32, 64, (3, 3, 3), 32, 64),
    ("c128_64_w64", 1, 2, 6, 64, 128, 64, (3, 3, 3), 128, 64),
    ("c64_128_w32", 2, 2, 8, 32, 64, 128, (3, 3, 3), 64, 128),
    ("c128_128_w32", 1, 2, 8, 32, 128, 128, (3, 3, 3), 128, 128),
    ("pm12_12_w256", 1, 3, 9, 256, 12, 12, (3, 7, 7), 16, 16),
    ("final16_3", 1, 2, 4, 256, 16, 3, (1, 1, 1), 16, 3),
    ("odd_w24", 1, 3, 16, 24, 16, 16, (3, 3, 3), 16, 16),
    ("odd_w100_h7", 2, 2, 7, 100, 32, 32, (3, 3, 3), 32, 32),
    ("slice_in_cat", 1, 2, 4, 128, 16, 16, (3, 3, 3), 32, 16),
]


def run(name, B, T, H, W, Cin, Cout, ks, x_ld, y_ld, timed=False):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.zeros(B, T, H, W, x_ld, device=dev, dtype=torch.bfloat16)
    x[..., :Cin] = torch.randn(B, T, H, W, Cin, device=dev, generator=g).bfloat16()
    w = (torch.randn(*ks, Cin, Cout, device=dev, generator=g) * 0.2).bfloat16()
    bias = torch.randn(Cout, device=dev, generator=g)
    res = {"name": name}
    outs = {}
    for backend in (_ffi.BACKEND_SIMT, _ffi.BACKEND_AUTO):
        ops.CONV_BACKEND = backend
        wp = ops.conv3d_wprep(w, 0, B, T, H, W, Cin, Cout, ks, x_ld, y_ld) if backend != _ffi.BACKEND_SIMT else None
        if backend != _ffi.BACKEND_SIMT:
            res["tc_fwd"] = wp is not None
        y = torch.zeros(B, T, H, W, y_ld, device=dev, dtype=torch.bfloat16)
        ops.conv3d_fwd(x, w, bias, ks, Cin, Cout, x_ld=x_ld, out=y, out_ld=y_ld, wprep=wp)
        torch.cuda.synchronize()
        outs[("fwd", backend)] = y.float()
        # dgrad: dy has Cout channels (stride y_ld), dx gets Cin channels (stride x_ld)
        dy = torch.zeros(B, T, H, W, y_ld, device=dev, dtype=torch.bfloat16)
        dy[..., :Cout] = torch.randn(B, T, H, W, Cout, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).bfloat16()
        wpd = ops.conv3d_wprep(w, 1, B, T, H, W, Cin, Cout, ks, x_ld, y_ld) if backend != _ffi.BACKEND_SIMT else None
        if backend != _ffi.BACKEND_SIMT:
            res["tc_dgrad"] = wpd is not None
        dx = torch.zeros(B, T, H, W, x_ld, device=dev, dtype=torch.bfloat16)
        ops.conv3d_dgrad(dy, w, ks, Cin, Cout, dy_ld=y_ld, out=dx, out_ld=x_ld, wprep=wpd)
        torch.cuda.synchronize()
        outs[("dgrad", backend)] = dx.float()
        # wgrad: dw += x^T (*) dy
        dwt = torch.zeros(*ks, Cin, Cout, device=dev, dtype=torch.float32)
        ops.conv3d_wgrad_accum(x, dy, dwt, ks, Cin, Cout, x_ld=x_ld, dy_ld=y_ld)
        torch.cuda.synchronize()
        outs[("wgrad", backend)] = dwt.clone()
        if timed and backend != _ffi.BACKEND_SIMT:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv3d_wgrad_accum(x, dy, dwt, ks, Cin, Cout, x_ld=x_ld, dy_ld=y_ld)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            res["wgrad_ms"] = ms
            res["wgrad_tflops"] = 2.0 * B * T * H * W * ks[0] * ks[1] * ks[2] * Cin * Cout / ms / 1e9
        if timed and backend != _ffi.BACKEND_SIMT and wp is not None:
            for kind in ("fwd", "dgrad"):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    if kind == "fwd":
                        ops.conv3d_fwd(x, w, bias, ks, Cin, Cout, x_ld=x_ld, out=y, out_ld=y_ld, wprep=wp)
                    else:
                        ops.conv3d_dgrad(dy, w, ks, Cin, Cout, dy_ld=y_ld, out=dx, out_ld=x_ld, wprep=wpd)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                fl = 2.0 * B * T * H * W * ks[0] * ks[1] * ks[2] * Cin * Cout
                res[kind + "_ms"] = ms
                res[kind + "_tflops"] = fl / ms / 1e9
    for kind in ("fwd", "dgrad", "wgrad"):
        a, b = outs[(kind, _ffi.BACKEND_AUTO)], outs[(kind, _ffi.BACKEND_SIMT)]
        res[kind + "_err"] = ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()
    return res


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    sel = sys.argv[1:]
    with open("gpurun_out/conv_bringup.jsonl", "a") as f:
        for c in CASES:
            if sel and c[0] not in sel:
                continue
            try:
                r = run(*c)
            except Exception as e:  # noqa: BLE001
                r = {"name": c[0], "error": repr(e)[:300]}
            print(json.dumps(r), flush=True)
            f.write(json.dumps(r) + "\n")
        if True:
            for c in [("prod_full16", 8, 16, 256, 256, 16, 16, (3, 3, 3), 16, 16),
                      ("prod_full32_16", 8, 16, 256, 256, 32, 16, (3, 3, 3), 32, 16),
                      ("prod_pm", 8, 16, 256, 256, 12, 12, (3, 7, 7), 16, 16),
                      ("prod_half64_32", 8, 16, 128, 128, 64, 32, (3, 3, 3), 64, 32),
                      ("prod_q128_64", 8, 16, 64, 64, 128, 64, (3, 3, 3), 128, 64),
                      ("prod_b128_128", 8, 16, 32, 32, 128, 128, (3, 3, 3), 128, 128)]:
                if sel and c[0] not in sel:
                    continue
                try:
                    r = run(*c, timed=True)
                except Exception as e:  # noqa: BLE001
                    r = {"name": c[0], "error": repr(e)[:300]}
                print(json.dumps(r), flush=True)
                f.write(json.dumps(r) + "\n")
