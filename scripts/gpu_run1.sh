set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 5 --warmup 3 --profile-kernels > gpurun_out/bench3.json 2> gpurun_out/bench3_classes.log; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
