set -x
python tests/gpu_bringup_attn.py prod_temporal > gpurun_out/plain_attn_warp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_warp_bwd -s 2 -c 1 -f -o gpurun_out/prof_attn_warp_bwd python tests/gpu_bringup_attn.py prod_temporal > gpurun_out/ncu_attn_warp.log 2>&1
echo ncu rc=$?
