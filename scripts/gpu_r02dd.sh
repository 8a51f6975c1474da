#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ln_one.py && timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm_ -c 6 -o gpurun_out/r02dd_ln -f python scripts/ln_one.py > gpurun_out/r02dd_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02dd_ncu.log
