# A/B of BUILD flags: CONFIGS="flags1|flags2" (| separated), e.g. CONFIGS="-DFL_THREADS=128 -DFL_MINB=7|-DFL_THREADS=256 -DFL_MINB=3"
set -x
mkdir -p gpurun_out/abb
IFS='|' read -ra CFG <<< "$CONFIGS"
i=0
for F in "${CFG[@]}"; do
  SPHE_NVCC_EXTRA="$F" python sph-erosion_b200/build.py > gpurun_out/abb/build_$i.log 2>&1 || tail -5 gpurun_out/abb/build_$i.log
  for W in ${WL:-c3}; do
    timeout 600 python bench.py --workload $W --no-cpu-baseline --no-reference-gravity --no-parity-gate ${BARGS} > gpurun_out/abb/bench_${W}_$i.json 2> gpurun_out/abb/bench_${W}_$i.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/abb/bench_${W}_$i.json"))
    print("ABB [$F] $W", round(d["ms_per_step"],4), "%.3e"%d["value"], {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms_per_step"].items()})
except Exception as e: print("$W failed", e)
PY
  done
  i=$((i+1))
done
