# last GPU session of round 2: smoke, the whole GPU suite, the default bench line, c2 line, density capture with the 256-bit loads
set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_c3_default_n1.json 2> $O/bench_c3.err; echo "bench rc=$?"
timeout 900 python bench.py --workload c2 > $O/bench_c2_n1.json 2> $O/bench_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
ARGS="--steps 6 --warmup 3 --no-cpu-baseline --no-parity-gate --no-reference-gravity"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_density_list -s 155 -c 1 -f -o $O/prof_k_density_list_c3 python bench.py $ARGS > $O/ncu_density.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c3_all.csv python bench.py $ARGS > $O/launches_c3.log 2>&1
python - <<PY
import json
for f in ("bench_c3_default_n1","bench_c2_n1","bench_reference_arm"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), "%.4e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d.get("parity_sampled"), d.get("gpu_launches"))
    except Exception as e: print(f, "failed", e)
PY
