# A/B of kernel variants on c2 and c3-noterrain
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for W in c2 c3-noterrain; do
for v in ${VARIANTS:-3 52 54 58}; do
  timeout 600 python bench.py --workload $W --steps 100 --warmup 5 --no-cpu-baseline --density-variant $v --force-variant $v > gpurun_out/ab_${W}_v$v.json 2> gpurun_out/ab_${W}_v$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_${W}_v$v.json'))
print('$W variant $v', 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], {k:round(x,4) for k,x in d['roofline']['per_kernel_ms_per_step'].items()})
PY
done; done
