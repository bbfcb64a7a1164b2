set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 72 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
tail -2 gpurun_out/ncu_bench3.log
timeout 600 python -m pytest tests/test_gpu_terrain.py -m gpu -x -q --durations=8 2>&1 | tail -14
