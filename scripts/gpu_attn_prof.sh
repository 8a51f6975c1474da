set -x
python tests/gpu_bringup_attn.py prod_spatial > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd python tests/gpu_bringup_attn.py prod_spatial > gpurun_out/ncu_attn.log 2>&1
echo ncu rc=$?
