#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_prod_gpu.py -q -x -m gpu > gpurun_out/r02ee_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02ee_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-kernels > gpurun_out/r02ee_bench.json 2> gpurun_out/r02ee_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02ee_bench.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved'])"
grep abi_entry gpurun_out/r02ee_bench.err | head -14
