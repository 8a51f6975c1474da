#!/usr/bin/env python
"""Per-layer timing of the U-Net's conv3d kernels at BASELINE configs[1] shapes.

    python scripts/conv_probe.py            -> one JSON line per layer: fwd / dgrad / wgrad ms and TFLOP/s (CUDA events)
    python scripts/conv_probe.py --ncu      -> launches only the three shapes VERDICT r1 asked ncu captures for
                                               (patch_mixer 3x7x7 12->12, 3x3x3 16->16 fwd, and the 16->16 wgrad)

Each timed launch works on inputs far larger than L2 for the full-resolution layers (201-268 MB per map).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402

B, T = 8, 16
LAYERS = [
    # name, H(=W), Cin, Cout, ks, x_ld, y_ld
    ("patch_mixer", 256, 12, 12, (3, 7, 7), 16, 16),
    ("enc0.conv1", 256, 12, 16, (3, 3, 3), 16, 16),
    ("enc0.conv2", 256, 16, 16, (3, 3, 3), 16, 16),
    ("enc1.conv1", 128, 16, 32, (3, 3, 3), 16, 32),
    ("enc1.conv2", 128, 32, 32, (3, 3, 3), 32, 32),
    ("enc2.conv1", 64, 32, 64, (3, 3, 3), 32, 64),
    ("enc2.conv2", 64, 64, 64, (3, 3, 3), 64, 64),
    ("bottleneck1", 32, 64, 128, (3, 3, 3), 64, 128),
    ("bottleneck2", 32, 128, 128, (3, 3, 3), 128, 128),
    ("dec0.conv1", 64, 128, 64, (3, 3, 3), 128, 64),
    ("dec0.conv2", 64, 64, 64, (3, 3, 3), 64, 64),
    ("dec1.conv1", 128, 64, 32, (3, 3, 3), 64, 32),
    ("dec1.conv2", 128, 32, 32, (3, 3, 3), 32, 32),
    ("dec2.conv1", 256, 32, 16, (3, 3, 3), 32, 16),
    ("dec2.conv2", 256, 16, 16, (3, 3, 3), 16, 16),
]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def setup(H, Cin, Cout, ks, x_ld, y_ld):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.zeros(B, T, H, H, x_ld, device=dev, dtype=torch.bfloat16)
    x[..., :Cin] = torch.randn(B, T, H, H, Cin, device=dev, generator=g).bfloat16()
    dy = torch.zeros(B, T, H, H, y_ld, device=dev, dtype=torch.bfloat16)
    dy[..., :Cout] = torch.randn(B, T, H, H, Cout, device=dev, generator=g).bfloat16()
    w = (torch.randn(*ks, Cin, Cout, device=dev, generator=g) * 0.1).bfloat16()
    bias = torch.randn(Cout, device=dev, generator=g)
    return x, dy, w, bias


def run_layer(name, H, Cin, Cout, ks, x_ld, y_ld, which=("fwd", "dgrad", "wgrad"), reps=3):
    x, dy, w, bias = setup(H, Cin, Cout, ks, x_ld, y_ld)
    flop = 2.0 * B * T * H * H * ks[0] * ks[1] * ks[2] * Cin * Cout
    res = {"layer": name, "voxels": B * T * H * H, "Cin": Cin, "Cout": Cout, "ks": list(ks), "gflop": flop / 1e9}
    y = torch.zeros(B, T, H, H, y_ld, device="cuda", dtype=torch.bfloat16)
    dx = torch.zeros(B, T, H, H, x_ld, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(*ks, Cin, Cout, device="cuda", dtype=torch.float32)
    wp = ops.conv3d_wprep(w, 0, B, T, H, H, Cin, Cout, ks, x_ld, y_ld)
    wpd = ops.conv3d_wprep(w, 1, B, T, H, H, Cin, Cout, ks, x_ld, y_ld)
    res["tensor_core_path"] = wp is not None and wpd is not None
    fns = {
        "fwd": lambda: ops.conv3d_fwd(x, w, bias, ks, Cin, Cout, x_ld=x_ld, out=y, out_ld=y_ld, wprep=wp),
        "dgrad": lambda: ops.conv3d_dgrad(dy, w, ks, Cin, Cout, dy_ld=y_ld, out=dx, out_ld=x_ld, wprep=wpd),
        "wgrad": lambda: ops.conv3d_wgrad_accum(x, dy, dw, ks, Cin, Cout, x_ld=x_ld, dy_ld=y_ld),
    }
    io_bytes = {"fwd": B * T * H * H * (x_ld + y_ld) * 2, "dgrad": B * T * H * H * (x_ld + y_ld) * 2,
                "wgrad": B * T * H * H * (x_ld + y_ld) * 2}
    for k in which:
        ms = timed(fns[k], reps)
        res[k + "_ms"] = round(ms, 4)
        res[k + "_tflops"] = round(flop / ms / 1e9, 1)
        res[k + "_io_gbs"] = round(io_bytes[k] / ms / 1e6, 1)
    return res


def main():
    _ffi.require_device()
    if "--ncu" in sys.argv:
        for (name, which) in (("patch_mixer", ("fwd", "wgrad")), ("enc0.conv2", ("fwd", "dgrad", "wgrad"))):
            layer = next(l for l in LAYERS if l[0] == name)
            print(json.dumps(run_layer(*layer, which=which, reps=1)), flush=True)
        return
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for layer in LAYERS:
        r = run_layer(*layer)
        for k in tot:
            tot[k] += r[k + "_ms"]
        print(json.dumps(r), flush=True)
    print(json.dumps({"layer": "TOTAL (15 conv3d layers; convT / 1x1x1 excluded)", **{k + "_ms": round(v, 3) for k, v in tot.items()},
                      "all_ms": round(sum(tot.values()), 3)}), flush=True)


if __name__ == "__main__":
    main()
