# the driver's scaling sequence: default bench at N = 1, 2, 4, 8 back to back (+ the reference arm once)
set -x
mkdir -p gpurun_out/scale
O=gpurun_out/scale
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
for N in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.load(open("gpurun_out/scale/bench_n%d.json"%n))
        if n==1: base=d["value"]
        print("N=%d ms/step %.4f value %.4e eff %.3f e2e %.3e parity %s" % (n, d["ms_per_step"], d["value"], d["value"]/(n*base), d["e2e"]["value"], d["parity_sampled"]))
    except Exception as e: print(n, "failed", e)
r=json.load(open("gpurun_out/scale/bench_reference.json")); print("reference", r["value"], r["extrapolated_full_n_value"])
PY
