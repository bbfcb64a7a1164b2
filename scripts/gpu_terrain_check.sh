set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_terrain.py tests/test_gpu_parity.py tests/test_host_shim.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --workload c3 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print('c3 ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], {k:round(x,4) for k,x in d['roofline']['per_kernel_ms_per_step'].items()}, d['config']['terrain_contacts_per_step'], d['config']['conservation_exact'])
PY
