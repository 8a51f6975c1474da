#!/bin/bash
# after the persistent attention forward: GPU tests, bench (A/B against the one-tile-per-CTA kernel), ncu of the new kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02ab_pytest.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --debug-set 18=1 > gpurun_out/r02ab_bench_oldfwd.json 2> gpurun_out/r02ab_bench_oldfwd.err; echo "bench(old fwd) rc=$?"
python - <<'PY'
import json
for f in ("r02ab_bench", "r02ab_bench_oldfwd"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, round(d["value"], 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 2), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"], 3))
    for k in d.get("kernel_classes", []):
        if "attn" in k["kernel"]: print("   ", k["kernel"], round(k["ms_per_step"], 3), k.get("tflops"))
PY
timeout 200 python scripts/attn_one.py > gpurun_out/r02ab_attn_one.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd256 -s 1 -c 1 -f -o gpurun_out/r02ab_attn_fwd256 python scripts/attn_one.py > gpurun_out/r02ab_ncu.log 2>&1
echo "ncu rc=$?"
