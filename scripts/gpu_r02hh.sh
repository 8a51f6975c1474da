#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r02hh_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02hh_pytest.log
for v in "" "--debug-set 17=1" ""; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v > gpurun_out/r02hh_bench.json 2> gpurun_out/r02hh_bench.err; echo "bench [$v] rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02hh_bench.json')); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
done
