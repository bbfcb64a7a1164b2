set -x
mkdir -p gpurun_out/r02f
O=gpurun_out/r02f
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -8 $O/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --no-reference-gravity > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; tail -c 800 $O/bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02f/bench_default.json"))
    print(d["config"]["workload"], d["ms_per_step"], d["value"], d["roofline"]["per_kernel_ms_per_step"], d["e2e"]["value"])
    print("parity", d["parity_sampled"], d["parity_gate"]["height_max_diff_fx"], d["config"]["conservation_exact"], d["config"]["terrain_contacts_per_step"])
except Exception as e: print("failed", e)
PY
