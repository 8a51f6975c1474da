#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/attn_one.py && timeout 300 python scripts/norm_one.py && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd256|attn_fwd_sm100|attn_bwd_sm100|attn_delta' -s 4 -c 4 -o gpurun_out/r02pp_attn -f python scripts/attn_one.py > gpurun_out/r02pp_ncu_attn.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02pp_ncu_attn.log
