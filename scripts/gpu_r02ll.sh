#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r02ll_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ll_pytest.log
for v in "" "--no-wgrad-lane" "" "--no-wgrad-lane"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v > gpurun_out/r02ll_bench.json 2> gpurun_out/r02ll_bench.err; echo "bench [$v] rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02ll_bench.json')); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'], d['peak_mem_GB'])"
done
