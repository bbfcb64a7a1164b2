set -x
mkdir -p gpurun_out/r02e
O=gpurun_out/r02e
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/bench_c3_n$N.json 2> $O/bench_c3_n$N.err; echo "rc=$?"; tail -c 1500 $O/bench_c3_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02e/bench_c3_n$N.json"))
    print(d["config"]["workload"], d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"])
    print({k:d["config"].get(k) for k in ("particles","particles_conserved","conservation_exact","terrain_boundary_rows_identical","terrain_window_violations","terrain_contacts_per_step")})
    print(d["roofline"]["per_rank"])
except Exception as e: print("failed", e)
PY
