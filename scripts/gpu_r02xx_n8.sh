#!/bin/bash
# 8 GPUs: headline config after all round-2 kernel changes (regression + record)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r02xx_cfg2_n8.json 2> gpurun_out/r02xx_cfg2_n8.err; echo "n8 rc=$?"
python -c "
import json
for l in open('gpurun_out/r02xx_cfg2_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n8', round(d['value'],2), d['unit'], round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],2), d['clocks'], d.get('grad_comm'))"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02xx_cfg2_n1.json 2> gpurun_out/r02xx_cfg2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02xx_cfg2_n1.json')); print('n1', round(d['value'],2), round(d['ms_per_step'],2), d['clocks'])"
