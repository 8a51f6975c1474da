#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "recompute or graph" 2>&1 | tail -3
for c in cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_bench_$c.json 2> gpurun_out/r02s_bench_$c.err; echo "$c rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_$c.json')); print('$c', round(d['value'],2), d['unit'], round(d['ms_per_step'],2),'ms', 'e2e', round(d['e2e']['value'],2), 'mem', round(d['peak_mem_GB'],1), 'tflops', d['model_tflops'])"
done
python bench.py --config cfg5 --batch 2 --recompute --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02s_bench_cfg5_b2_recompute.json 2> gpurun_out/r02s_bench_cfg5_b2.err; echo "cfg5 b2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_cfg5_b2_recompute.json')); print('cfg5 b2 recompute', round(d['value'],2), round(d['ms_per_step'],2),'ms', 'mem', round(d['peak_mem_GB'],1))"
python bench.py --config cfg5 --batch 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02s_bench_cfg5_b2.json 2>> gpurun_out/r02s_bench_cfg5_b2.err; echo "cfg5 b2 plain rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_cfg5_b2.json')); print('cfg5 b2 stored', round(d['value'],2), round(d['ms_per_step'],2),'ms', 'mem', round(d['peak_mem_GB'],1))"
