#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02t_pytest.log
for f in "" "--no-pdl"; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline $f > gpurun_out/r02t_bench$f.json 2> gpurun_out/r02t_bench$f.err; echo "bench $f rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02t_bench$f.json')); print('$f', d['value'], d['ms_per_step'], d['clocks'])"
done
