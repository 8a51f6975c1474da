#!/bin/bash
mkdir -p gpurun_out
python scripts/conv_probe.py --ncu > gpurun_out/r02l_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_sm100_kernel -s 2 -c 1 -o gpurun_out/r02l_conv python scripts/conv_probe.py --ncu > gpurun_out/r02l_ncu.log 2>&1; echo "ncu rc=$?"
