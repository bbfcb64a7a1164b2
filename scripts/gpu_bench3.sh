set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload c3 --steps 100 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -5 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
timeout 600 python bench.py --workload c3-noterrain --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3nt.json 2> gpurun_out/bench_c3nt.err; tail -5 gpurun_out/bench_c3nt.err; cat gpurun_out/bench_c3nt.json
