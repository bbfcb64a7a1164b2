# ONE long dam break over N GPUs WITH the eroding terrain (c3, --layout contiguous), without and with re-cuts:
# the slab boundaries and the terrain row windows move together (TerrainWindowShare.recut)
set -x
mkdir -p gpurun_out/imb_t
O=gpurun_out/imb_t
N=${N:-4}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@"; }
for RB in 0 25; do
run --workload c3 --layout contiguous --steps 200 --warmup 10 --settle 150 --rebalance-every $RB > $O/bench_c3_dam_n${N}_rb$RB.json 2> $O/bench_c3_dam_n${N}_rb$RB.err; echo "rc=$?"; tail -c 400 $O/bench_c3_dam_n${N}_rb$RB.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_c3_dam_n${N}_rb$RB.json").read().strip().splitlines()[-1])
    c=d["config"]
    print("rebalance-every $RB:", d["n_gpus"], round(d["ms_per_step"],4), "%.3e"%d["value"], "recuts", c.get("recuts_done"), [(r["particles"], r["owned"]) for r in d["roofline"]["per_rank"]], c.get("slab_columns"), c.get("terrain_window_rows"))
    print("  parity", d["parity_sampled"], "conservation", c.get("conservation_exact"), "violations", c.get("terrain_window_violations"), "boundary rows identical", c.get("terrain_boundary_rows_identical"))
    print("  gate", json.dumps(d["parity_gate"])[:600])
except Exception as e: print("failed", e)
PY
done
