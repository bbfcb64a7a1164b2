set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for v in "0 0" "1 0" "1 1"; do
  set -- $v
  timeout 600 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu-baseline --density-variant $1 --force-variant $2 > gpurun_out/bench_c2_v$1$2.json 2> gpurun_out/bench_c2_v$1$2.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_v$1$2.json'))
print('variant $1 $2', 'ms/step', d['ms_per_step'], 'value %.3e'%d['value'], d['roofline']['per_kernel_ms_per_step'])
PY
done
timeout 600 python bench.py --workload c3 --steps 50 --warmup 5 --no-cpu-baseline --density-variant 1 --force-variant 1 > gpurun_out/bench_c3_v11.json 2> gpurun_out/bench_c3_v11.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c3_v11.json'))
print('c3 v11 ms/step', d['ms_per_step'], 'value %.3e'%d['value'], d['roofline']['per_kernel_ms_per_step'])"
