# c2 (1M particles per GPU, no terrain, one long dam break: --layout contiguous by default) at N = 1, 2, 4, 8
set -x
mkdir -p gpurun_out/scale_c2
O=gpurun_out/scale_c2
timeout 600 python bench.py --workload c2 > $O/bench_n1.json 2> $O/bench_n1.err
for N in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --workload c2 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/scale_c2/bench_n%d.json"%n).read().strip().splitlines()[-1])
        if n==1: base=d["value"]
        print("N=%d ms/step %.4f value %.4e eff %.3f e2e %.3e parity %s" % (n, d["ms_per_step"], d["value"], d["value"]/(n*base), d["e2e"]["value"], d["parity_sampled"]))
    except Exception as e: print(n, "failed", e)
PY
