#!/bin/bash
# 8 GPUs: gradient all-reduce variants (one all-reduce after the graph / decoder slice overlapped with the encoder backward)
mkdir -p gpurun_out
run() { # name, env, args...
  name=$1; envs=$2; shift; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@" \
      > gpurun_out/r02yy_$name.json 2> gpurun_out/r02yy_$name.err; echo "$name rc=$?"
  python -c "
import json
for l in open('gpurun_out/r02yy_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['value'],2), d['unit'], round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d.get('grad_comm'))"
}
run plain "X=1" --steps 10 --warmup 3 --no-cpu-baseline
run split "X=1" --steps 10 --warmup 3 --no-cpu-baseline --split-allreduce
run split_cta16 "NCCL_MAX_CTAS=16" --steps 10 --warmup 3 --no-cpu-baseline --split-allreduce
run plain2 "X=1" --steps 10 --warmup 3 --no-cpu-baseline
