#!/bin/bash
# round 2, call A: full GPU test suite (incl. the new production-depth parity tests), headline bench, conv per-layer table,
# ncu --set full of the conv kernels VERDICT r1 asked for
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02a_pytest.log
python bench.py --steps 5 --warmup 3 --profile-kernels > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
python scripts/conv_probe.py > gpurun_out/r02a_conv_probe.jsonl 2> gpurun_out/r02a_conv_probe.err; echo "probe rc=$?"
tail -3 gpurun_out/r02a_conv_probe.jsonl
python scripts/conv_probe.py --ncu > gpurun_out/r02a_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_sm100_kernel|conv_wgrad_sm100_kernel' -c 10 \
    -o gpurun_out/r02a_conv python scripts/conv_probe.py --ncu > gpurun_out/r02a_ncu.log 2>&1; echo "ncu rc=$?"
