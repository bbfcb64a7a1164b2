# Round-1 re-entry baseline: GPU tests, smoke, bench c2/c3, ncu launch list.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 600 python bench.py --workload c2 --steps 100 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c3 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
