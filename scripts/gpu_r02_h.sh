set -x
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
N=${N:-2}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N "$@"; }
run > $O/bench_c3_n$N.json 2> $O/bench_c3_n$N.err; echo "rc=$?"; tail -c 600 $O/bench_c3_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02h/bench_c3_n$N.json"))
    print(d["config"]["workload"], d["n_gpus"], d["ms_per_step"], "%.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"])
    print("parity", d["parity_sampled"], json.dumps(d["parity_gate"]))
except Exception as e: print("failed", e)
PY
if [ "$IMB" = "1" ]; then
for RB in 0 25; do
run --workload c2 --steps 200 --warmup 300 --rebalance-every $RB --no-parity-gate > $O/bench_c2_dam_n${N}_rb$RB.json 2> $O/bench_c2_dam_n${N}_rb$RB.err; tail -c 300 $O/bench_c2_dam_n${N}_rb$RB.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02h/bench_c2_dam_n${N}_rb$RB.json"))
    print("rebalance-every $RB:", d["n_gpus"], round(d["ms_per_step"],4), "%.3e"%d["value"], [ (r["particles"], r["owned"]) for r in d["roofline"]["per_rank"]], d["config"].get("slab_columns"))
except Exception as e: print("failed", e)
PY
done
fi
