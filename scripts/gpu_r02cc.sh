#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/norm_ab.py > gpurun_out/r02cc_norm_ab.json 2> gpurun_out/r02cc_norm_ab.err; echo "ab rc=$?"; cat gpurun_out/r02cc_norm_ab.json; tail -5 gpurun_out/r02cc_norm_ab.err
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "layernorm or factored or videovae" > gpurun_out/r02cc_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02cc_pytest.log
