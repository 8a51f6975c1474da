#!/bin/bash
# launch list of the bench command (contract: ncu --metrics gpu__time_duration.sum --clock-control none of the same command)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02bb_plain.json 2> gpurun_out/r02bb_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 3200 --csv --log-file gpurun_out/r02bb_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02bb_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02bb_launches.csv
