set -x
mkdir -p gpurun_out/r02k
O=gpurun_out/r02k
timeout 900 python -m pytest tests/test_gpu_multiproc.py tests/test_gpu_slabs.py -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -6 $O/pytest.log
for W in 3 8; do
  SPHE_ONE_GPU=1 STEPS=12 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29500+W)) scripts/peer_check.py > $O/peer_check_one_gpu_w$W.log 2>&1
  echo "rc=$?" >> $O/peer_check_one_gpu_w$W.log; grep -E "PEER_CHECK|rc=" $O/peer_check_one_gpu_w$W.log
done
