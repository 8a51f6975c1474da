#!/usr/bin/env python
"""Timing ablations of the tcgen05 conv kernel on one full-resolution 16->16 layer (vvae_debug_set(14, bits))."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops
from scripts.conv_probe import LAYERS, run_layer
_ffi.require_device()
names = sys.argv[1:] or ["enc0.conv2", "patch_mixer", "dec2.conv1"]
for name in names:
    layer = next(l for l in LAYERS if l[0] == name)
    for bits, what in ((0, "full"), (1, "no stores"), (2, "no epilogue"), (4, "no input TMA"), (8, "no MMA"), (6, "no epilogue + no TMA"),
                       (10, "no epilogue + no MMA"), (12, "no TMA + no MMA"), (14, "only weights + barriers")):
        _ffi.lib.vvae_debug_set(14, bits)
        r = run_layer(*layer, which=("fwd", "dgrad"), reps=3)
        print(json.dumps({"layer": name, "ablation": what, "fwd_ms": r["fwd_ms"], "dgrad_ms": r["dgrad_ms"]}), flush=True)
    _ffi.lib.vvae_debug_set(14, 0)
    _ffi.lib.vvae_debug_set(13, 1)
    r = run_layer(*layer, which=("fwd", "dgrad"), reps=3)
    print(json.dumps({"layer": name, "ablation": "round-1 kernel (one MMA per tap)", "fwd_ms": r["fwd_ms"], "dgrad_ms": r["dgrad_ms"]}), flush=True)
    _ffi.lib.vvae_debug_set(13, 0)
