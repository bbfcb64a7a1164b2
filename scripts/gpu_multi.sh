# usage: N=<gpus> W=<workload> bash scripts/gpu_multi.sh
set -x
mkdir -p gpurun_out
N=${N:-2}; W=${W:-c2}
if [ -n "$PYTEST" ]; then timeout 900 python -m pytest $PYTEST -x -q 2>&1 | tail -15; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps ${STEPS:-50} --warmup 5 > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err
tail -5 gpurun_out/bench_${W}_n$N.err; cat gpurun_out/bench_${W}_n$N.json
