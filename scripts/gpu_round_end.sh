#!/bin/bash
# what the driver runs at round end, plus the launch list: GPU tests, smoke, reference arm, our arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/re_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/re_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/re_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/re_smoke.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/re_ref.json 2> gpurun_out/re_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/re_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/re_bench.json 2> gpurun_out/re_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/re_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['cpu_baseline'])"
