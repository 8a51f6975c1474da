#!/bin/bash
# end of round 2 (session 4): what the driver runs (GPU tests, smoke, reference arm, our arm) + the launch list of the bench command
mkdir -p gpurun_out
bash scripts/gpu_round_end.sh
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02zz_plain.json 2> gpurun_out/r02zz_plain.err && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 5400 -c 3600 --csv --log-file gpurun_out/r02zz_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02zz_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02zz_launches.csv
python scripts/summarize_launches.py gpurun_out/r02zz_launches.csv --steps-between adam_vec4_kernel 2 > gpurun_out/r02zz_launches.md; head -24 gpurun_out/r02zz_launches.md
