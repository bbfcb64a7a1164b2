# A/B of environment settings: ENVS="A=1|A=2" WL=c3
set -x
mkdir -p gpurun_out/abe
IFS='|' read -ra CFG <<< "$ENVS"
i=0
for E in "${CFG[@]}"; do
  for W in ${WL:-c3}; do
    env $E timeout 600 python bench.py --workload $W --no-cpu-baseline --no-reference-gravity --no-parity-gate > gpurun_out/abe/bench_${W}_$i.json 2> gpurun_out/abe/bench_${W}_$i.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/abe/bench_${W}_$i.json"))
    print("ABE [$E] $W", round(d["ms_per_step"],4), "%.3e"%d["value"], {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms_per_step"].items()})
except Exception as e: print("$W failed", e)
PY
  done
  i=$((i+1))
done
