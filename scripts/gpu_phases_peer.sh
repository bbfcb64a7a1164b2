set -x
mkdir -p gpurun_out
W=${W:-c3} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29513 scripts/slab_phases_peer.py 2>&1 | grep "^rank" | tee gpurun_out/slab_phases_peer_${W:-c3}_n${N:-2}.log
