set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 > gpurun_out/bench_c2_n4_default.json 2> gpurun_out/bench_c2_n4_default.err
tail -2 gpurun_out/bench_c2_n4_default.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err
cat gpurun_out/bench_ref_n4.json | cut -c1-300
