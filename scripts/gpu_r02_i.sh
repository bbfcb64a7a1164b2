set -x
mkdir -p gpurun_out/r02i
O=gpurun_out/r02i
N=${N:-4}
timeout 600 python -m pytest tests/test_gpu_slabs.py -q -x > $O/pytest_slabs.log 2>&1; echo "rc=$?" >> $O/pytest_slabs.log; tail -4 $O/pytest_slabs.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N "$@"; }
for RB in 0 25 100; do
run --workload c2 --steps 400 --warmup 400 --rebalance-every $RB --no-parity-gate > $O/bench_c2_dam_n${N}_rb$RB.json 2> $O/bench_c2_dam_n${N}_rb$RB.err; tail -c 300 $O/bench_c2_dam_n${N}_rb$RB.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02i/bench_c2_dam_n${N}_rb$RB.json"))
    print("rebalance-every $RB:", d["n_gpus"], round(d["ms_per_step"],4), "%.3e"%d["value"], [ (r["particles"], r["owned"]) for r in d["roofline"]["per_rank"]], d["config"].get("slab_columns"))
except Exception as e: print("failed", e)
PY
done
