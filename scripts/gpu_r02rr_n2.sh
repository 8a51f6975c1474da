#!/bin/bash
# 2 GPUs: regression check of the data-parallel path after the kernel changes of the second half of round 2
mkdir -p gpurun_out
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" \
      > gpurun_out/r02rr_$name.json 2> gpurun_out/r02rr_$name.err; echo "$name rc=$?"
  python -c "
import json
for l in open('gpurun_out/r02rr_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['value'],2), d['unit'], round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],2), d['clocks'], d.get('grad_comm'))"
}
run cfg2_n2 --steps 10 --warmup 3 --no-cpu-baseline
run cfg4_n2 --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_ddp_equivalence.py > gpurun_out/r02rr_ddp_equiv.log 2>&1; echo "ddp equiv rc=$?"; tail -3 gpurun_out/r02rr_ddp_equiv.log
