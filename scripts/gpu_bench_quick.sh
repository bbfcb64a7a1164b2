set -x
mkdir -p gpurun_out/quick
for W in ${WL:-c3 c2}; do
timeout 600 python bench.py --workload $W --no-cpu-baseline --no-reference-gravity ${ARGS} > gpurun_out/quick/bench_$W.json 2> gpurun_out/quick/bench_$W.err; tail -c 300 gpurun_out/quick/bench_$W.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/quick/bench_$W.json"))
    print("$W", round(d["ms_per_step"],4), "%.3e"%d["value"], {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms_per_step"].items()}, "e2e %.3e"%d["e2e"]["value"], "parity", d["parity_sampled"])
except Exception as e: print("$W failed", e)
PY
done
