#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/gpu_bringup_attn.py > gpurun_out/r02gg_attn_bringup.log 2>&1; echo "bringup rc=$?"; grep -c '"name"' gpurun_out/r02gg_attn_bringup.log; python - <<'P'
import json
for l in open('gpurun_out/r02gg_attn_bringup.log'):
    try: r=json.loads(l)
    except: continue
    bad = [k for k in ('o_err','dq_err','dk_err','dv_err') if r.get(k,0) > 0.02] or ('error' in r)
    if bad or r['name'].startswith('prod') or 'L256' in r['name']: print(r)
P
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_prod_gpu.py -q -x -m gpu > gpurun_out/r02gg_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02gg_pytest.log
for v in "" "--debug-set 16=1" ""; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v > gpurun_out/r02gg_bench.json 2> gpurun_out/r02gg_bench.err; echo "bench [$v] rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02gg_bench.json')); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
done
