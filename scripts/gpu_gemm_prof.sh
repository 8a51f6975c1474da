set -x
CASE=${1:-'{"M":32768,"N":768,"K":512,"tA":0,"tB":0,"bias":1,"mode":2,"time":1,"name":"prod_out_fwd"}'}
OUT=${2:-prof_gemm_out}
python tests/gpu_bringup_gemm.py --case "$CASE" > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_sm100 -s 2 -c 1 -f -o gpurun_out/$OUT python tests/gpu_bringup_gemm.py --case "$CASE" > gpurun_out/ncu_gemm.log 2>&1
echo ncu rc=$?
