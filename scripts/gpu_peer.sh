# usage: N=<gpus> bash scripts/gpu_peer.sh   -- multi-process peer exchange check + benches
set -x
mkdir -p gpurun_out
N=${N:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/peer_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -15 | tee gpurun_out/peer_check_n$N.log
for W in ${WORKLOADS:-c2 c3}; do
for X in ${EXCHANGES:-peer nccl}; do
  if [ "$W$X" = "c3nccl" ]; then continue; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --exchange $X --steps ${STEPS:-100} --warmup 5 > gpurun_out/bench_${W}_n${N}_$X.json 2> gpurun_out/bench_${W}_n${N}_$X.err
  tail -3 gpurun_out/bench_${W}_n${N}_$X.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${W}_n${N}_$X.json'))
    print('$W N=$N $X', 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(x,4) for k,x in d['roofline']['per_kernel_ms_per_step'].items()}, {k:v for k,v in d['config'].items() if k in ('conservation_exact','terrain_replicas_identical','terrain_contacts_per_step')})
except Exception as e: print('no result', e)
PY
done; done
