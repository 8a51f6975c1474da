#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/conv_bringup.jsonl
timeout 600 python tests/gpu_bringup_conv.py c16_16_w256 c12_16_w256_ld16 c32_16_w256 c16_32_w128 c64_32_w128 c32_64_w64 c128_64_w64 c64_128_w32 c128_128_w32 pm12_12_w256 final16_3 odd_w24 odd_w100_h7 slice_in_cat > gpurun_out/r02m_bringup.log 2>&1; echo "bringup rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r02m_bringup.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('name'), {k:(round(v,5) if isinstance(v,float) else v) for k,v in d.items() if k!='name'})
P
timeout 300 python scripts/conv_probe.py > gpurun_out/r02m_conv_probe.jsonl 2> gpurun_out/r02m_conv_probe.err; echo "probe rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r02m_conv_probe.jsonl'):
    d=json.loads(l)
    if 'voxels' in d: print(f"{d['layer']:14s} fwd {d['fwd_ms']:.3f} dgrad {d['dgrad_ms']:.3f} wgrad {d['wgrad_ms']:.3f}")
    else: print(d)
P
python scripts/conv_ablate.py enc0.conv2 2>&1 | tee gpurun_out/r02m_conv_ablate.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest.log
python bench.py --steps 5 --warmup 3 --profile-kernels > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for r in d['kernel_classes']: print(r)"
