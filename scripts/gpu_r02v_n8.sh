#!/bin/bash
# 8 GPUs: headline config with fp32 and bf16 gradient all-reduce, and BASELINE configs[4] (64x512x512, 1 clip per GPU)
mkdir -p gpurun_out
run() { # name, args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@" \
      > gpurun_out/r02v_$name.json 2> gpurun_out/r02v_$name.err; echo "$name rc=$?"
  python -c "
import json
for l in open('gpurun_out/r02v_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['value'],2), d['unit'], round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],2), d['clocks'], d.get('grad_comm'))"
}
run cfg2_fp32 --steps 10 --warmup 3 --no-cpu-baseline
run cfg2_bf16 --steps 10 --warmup 3 --no-cpu-baseline --grad-comm bf16
run cfg5_fp32 --config cfg5 --steps 5 --warmup 2 --no-cpu-baseline
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02v_cfg2_n1.json 2> gpurun_out/r02v_cfg2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02v_cfg2_n1.json')); print('n1', round(d['value'],2), round(d['ms_per_step'],2), d['clocks'])"
