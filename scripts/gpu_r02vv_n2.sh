#!/bin/bash
mkdir -p gpurun_out
run() { # name, env, args...
  name=$1; envs=$2; shift; shift
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-2} "$@" \
      > gpurun_out/r02vv_$name.json 2> gpurun_out/r02vv_$name.err; echo "$name rc=$?"
  python -c "
import json
for l in open('gpurun_out/r02vv_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['value'],2), d['unit'], round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d.get('grad_comm'), 'loss', d['e2e'].get('last_loss'))"
}
run plain "X=1" --steps 10 --warmup 3 --no-cpu-baseline
run split "X=1" --steps 10 --warmup 3 --no-cpu-baseline --split-allreduce
run split_cta8 "NCCL_MAX_CTAS=8" --steps 10 --warmup 3 --no-cpu-baseline --split-allreduce
run plain_cta8 "NCCL_MAX_CTAS=8" --steps 10 --warmup 3 --no-cpu-baseline
run plain2 "X=1" --steps 10 --warmup 3 --no-cpu-baseline
