set -x
mkdir -p gpurun_out
for W in ${WORKLOADS:-c2 c3}; do
  timeout 900 python bench.py --workload $W --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline > gpurun_out/bench_${W}_n1.json 2> gpurun_out/bench_${W}_n1.err
  tail -3 gpurun_out/bench_${W}_n1.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${W}_n1.json'))
print('$W N=1', 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(x,4) for k,x in d['roofline']['per_kernel_ms_per_step'].items()})
PY
done
