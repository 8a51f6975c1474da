#!/bin/bash
mkdir -p gpurun_out
for v in "" "--debug-set 7=16" "--debug-set 7=32" "--debug-set 7=17" "--debug-set 7=1" ""; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v > gpurun_out/r02ff_bench.json 2> gpurun_out/r02ff_bench.err; echo "bench [$v] rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02ff_bench.json')); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
done
