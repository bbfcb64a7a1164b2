set -x
for LAG in 0 2; do
LAG=$LAG timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29512 scripts/slab_phases.py 2>&1 | grep "^rank"
done
