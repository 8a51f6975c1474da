#!/bin/bash
# launch list of the bench command (contract: ncu --metrics gpu__time_duration.sum --clock-control none of the same command)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02jj_plain.json 2> gpurun_out/r02jj_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 5400 -c 3600 --csv --log-file gpurun_out/r02jj_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02jj_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02jj_launches.csv
python scripts/summarize_launches.py gpurun_out/r02jj_launches.csv --steps-between adam_kernel 2 > gpurun_out/r02jj_launches.md; head -50 gpurun_out/r02jj_launches.md
