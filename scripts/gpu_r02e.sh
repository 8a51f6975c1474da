#!/bin/bash
mkdir -p gpurun_out
python scripts/conv_probe.py --ncu > gpurun_out/r02e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_sm100_kernel' -c 4 -o gpurun_out/r02e_conv python scripts/conv_probe.py --ncu > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
