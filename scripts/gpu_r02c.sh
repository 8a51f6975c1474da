#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/conv_probe.py > gpurun_out/r02c_conv_probe.jsonl 2> gpurun_out/r02c_conv_probe.err; echo "probe rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r02c_conv_probe.jsonl'):
    d=json.loads(l)
    if 'voxels' in d: print(f"{d['layer']:14s} fwd {d['fwd_ms']:.3f} dgrad {d['dgrad_ms']:.3f} wgrad {d['wgrad_ms']:.3f}")
    else: print(d)
P
