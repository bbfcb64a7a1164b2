# round 2, call A: state of the tree at the start of the round + the multi-process checks the verdict asked for
set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
for W in 2 4 8; do
  SPHE_ONE_GPU=1 STEPS=12 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29500+W)) scripts/peer_check.py > $O/peer_check_one_gpu_w$W.log 2>&1
  echo "rc=$?" >> $O/peer_check_one_gpu_w$W.log; grep -E "PEER_CHECK|bit-equal|conservation|rc=" $O/peer_check_one_gpu_w$W.log
done
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 1 --log-file $O/sanitize_$TOOL.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_slabs.py -x -q \
    -k "default and (golden_step1 or step1) or one_cell and default or peer_mailbox and 3" > $O/sanitize_${TOOL}_pytest.log 2>&1
  echo "rc=$?" >> $O/sanitize_${TOOL}_pytest.log; tail -3 $O/sanitize_${TOOL}_pytest.log; tail -8 $O/sanitize_$TOOL.log
done
timeout 900 python bench.py --workload c3 --steps 50 --warmup 10 > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 600 $O/bench_c3.err
timeout 300 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
python - <<'PY'
import json
for w in ("c3","c2"):
    try:
        d=json.load(open("gpurun_out/r02a/bench_%s.json"%w)); print(w, d["ms_per_step"], d["value"], d["roofline"]["per_kernel_ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(w, "failed", e)
PY
