set -x
mkdir -p gpurun_out/r02d
O=gpurun_out/r02d
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; tail -c 1500 $O/bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02d/bench_default.json"))
    print(d["config"]["workload"], d["ms_per_step"], d["value"], d["roofline"]["per_kernel_ms_per_step"], d["e2e"]["value"])
    print("parity", d["parity_sampled"], json.dumps(d["parity_gate"]))
    print("refg", json.dumps(d["reference_gravity"]))
    print("roofline", d["roofline"]["frac"], d["roofline"]["whole_step"], d["roofline"]["per_kernel_hbm_frac"])
    print("terrain", {k:d["config"].get(k) for k in ("heightmap","terrain_contacts_per_step","conservation_exact","replayed_window_bit_identical","mean_neighbours_after_run")})
    print("cpu", d["cpu_baseline"])
except Exception as e: print("failed", e)
PY
