set -x
mkdir -p gpurun_out/r02c
O=gpurun_out/r02c
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -25 $O/pytest_gpu.log
timeout 300 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python - <<'PY'
import json
for w in ("c2",):
    try:
        d=json.load(open("gpurun_out/r02c/bench_%s.json"%w)); print(w, d["ms_per_step"], d["value"], d["roofline"]["per_kernel_ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(w, "failed", e)
PY
