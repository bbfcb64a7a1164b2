set -x
mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -25 $O/pytest_gpu.log
for W in 3 8; do
  SPHE_ONE_GPU=1 STEPS=12 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29500+W)) scripts/peer_check.py > $O/peer_check_one_gpu_w$W.log 2>&1
  echo "rc=$?" >> $O/peer_check_one_gpu_w$W.log; grep -E "PEER_CHECK|rc=" $O/peer_check_one_gpu_w$W.log
done
