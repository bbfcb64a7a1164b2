set -x
mkdir -p gpurun_out/ncu_r02
O=gpurun_out/ncu_r02
ARGS="--steps 6 --warmup 3 --settle 600 --no-cpu-baseline --no-parity-gate --no-reference-gravity"
timeout 900 python bench.py $ARGS > $O/bench_settle600.json 2> $O/bench_settle600.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/ncu_r02/bench_settle600.json")); print("settle600", d["ms_per_step"], d["roofline"]["per_kernel_ms_per_step"], d["config"]["terrain_contacts_per_step"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_terrain_contact -s 605 -c 1 -f -o $O/prof_k_terrain_contact_late python bench.py $ARGS > $O/ncu_terrain_late.log 2>&1
tail -1 $O/ncu_terrain_late.log
