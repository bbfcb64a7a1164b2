# A/B of kernel variants on the bench workloads: VARIANTS="6,3 12,3" WL="c3 c2"
set -x
mkdir -p gpurun_out/ab
for V in ${VARIANTS:-6,3 12,3}; do
DV=${V%,*}; FV=${V#*,}
for W in ${WL:-c3 c2}; do
timeout 600 python bench.py --workload $W --no-cpu-baseline --no-reference-gravity --no-parity-gate --density-variant $DV --force-variant $FV ${ARGS} > gpurun_out/ab/bench_${W}_${DV}_${FV}.json 2> gpurun_out/ab/bench_${W}_${DV}_${FV}.err; tail -c 300 gpurun_out/ab/bench_${W}_${DV}_${FV}.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab/bench_${W}_${DV}_${FV}.json"))
    print("AB $W ($DV,$FV)", round(d["ms_per_step"],4), "%.3e"%d["value"], {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms_per_step"].items()})
except Exception as e: print("$W failed", e)
PY
done
done
