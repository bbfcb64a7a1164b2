set -x
mkdir -p gpurun_out/numa
O=gpurun_out/numa
nvidia-smi topo -m > $O/topo.txt 2>&1
lscpu | grep -i "numa\|socket\|^CPU(s)" > $O/lscpu.txt
timeout 600 python bench.py --no-cpu-baseline --no-reference-gravity --no-parity-gate > $O/n1.json 2> $O/n1.err
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-parity-gate > $O/n$N.json 2> $O/n$N.err
python - <<PY
import json
for f in ("n1","n$N"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], "%.4e"%d["value"], "e2e %.4e"%d["e2e"]["value"], d["e2e"].get("pinned_buffers"))
    except Exception as e: print(f, "failed", e)
PY
cat $O/lscpu.txt; head -12 $O/topo.txt
