set -x
mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
N=${N:-8}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N "$@"; }
run > $O/bench_c3_n$N.json 2> $O/bench_c3_n$N.err; echo "rc=$?"; tail -c 400 $O/bench_c3_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02j/bench_c3_n$N.json"))
    print(d["config"]["workload"], d["n_gpus"], d["ms_per_step"], "%.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"])
    print("parity", d["parity_sampled"], json.dumps(d["parity_gate"]))
    print({k:d["config"].get(k) for k in ("particles","particles_conserved","conservation_exact","terrain_boundary_rows_identical","terrain_window_violations")})
except Exception as e: print("failed", e)
PY
N=$N W=c3 bash scripts/gpu_phases_peer.sh
for RB in 0 50; do
run --workload c2 --steps 400 --warmup 400 --rebalance-every $RB --no-parity-gate > $O/bench_c2_dam_n${N}_rb$RB.json 2> $O/bench_c2_dam_n${N}_rb$RB.err; tail -c 300 $O/bench_c2_dam_n${N}_rb$RB.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02j/bench_c2_dam_n${N}_rb$RB.json"))
    print("rebalance-every $RB:", d["n_gpus"], round(d["ms_per_step"],4), "%.3e"%d["value"], [ r["owned"] for r in d["roofline"]["per_rank"]])
except Exception as e: print("failed", e)
PY
done
