#!/bin/bash
# round 2, call B: packed-tap conv kernels (fwd/dgrad: kw taps in N; wgrad: kh taps in N) -- parity vs the generic kernels, timing
mkdir -p gpurun_out; rm -f gpurun_out/conv_bringup.jsonl
timeout 900 python tests/gpu_bringup_conv.py > gpurun_out/r02b_bringup.log 2>&1; echo "bringup rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r02b_bringup.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('name'), {k:(round(v,5) if isinstance(v,float) else v) for k,v in d.items() if k!='name'})
P
timeout 300 python scripts/conv_probe.py > gpurun_out/r02b_conv_probe.jsonl 2> gpurun_out/r02b_conv_probe.err; echo "probe rc=$?"
tail -1 gpurun_out/r02b_conv_probe.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_prod_gpu.py tests/test_jax_golden.py -m gpu -q -x -k "unet or videovae or vgg or prod or stand_in" > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02b_pytest.log
