#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "gemm or linear or qkv or mlp or factored" > gpurun_out/r02qq_pytest1.log 2>&1; echo "pytest gemm rc=$?"; tail -2 gpurun_out/r02qq_pytest1.log
timeout 300 python scripts/gemm_epi_ablate.py > gpurun_out/r02qq_gemm_epi.jsonl 2>&1; echo "epi rc=$?"; cat gpurun_out/r02qq_gemm_epi.jsonl | cut -c1-150
timeout 300 python scripts/gemm_tail_ab.py > gpurun_out/r02qq_gemm_tail.jsonl 2>&1; echo "tail rc=$?"; grep -c '"bit_identical": true' gpurun_out/r02qq_gemm_tail.jsonl; tail -1 gpurun_out/r02qq_gemm_tail.jsonl
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r02qq_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02qq_pytest.log
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"; done
