#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "lane or colsum" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02mm_bench_1gpu.json 2> gpurun_out/r02mm_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02mm_bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline'], d['e2e'], d['cpu_baseline'])"
timeout 120 python scripts/attn_timeline.py > gpurun_out/r02nn_attn_timeline.jsonl 2>&1; echo "timeline rc=$?"
timeout 300 python scripts/gemm_tail_ab.py > gpurun_out/r02oo_gemm_tail_ab.jsonl 2>&1; echo "tail rc=$?"; tail -1 gpurun_out/r02oo_gemm_tail_ab.jsonl
timeout 300 python scripts/norm_ab.py > gpurun_out/r02kk_norm_ab.jsonl 2>&1; echo "norm rc=$?"
