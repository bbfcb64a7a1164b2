# round 2, call B: first run of the staged kernels (variant 20)
set -x
mkdir -p gpurun_out/r02b
O=gpurun_out/r02b
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -k "default or stage_tpp" > $O/pytest_stage.log 2>&1; echo "rc=$?" >> $O/pytest_stage.log; tail -30 $O/pytest_stage.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file $O/sanitize_memcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -k "default and (golden_step1 or one_cell or odd_counts)" > $O/sanitize_pytest.log 2>&1; echo "rc=$?" >> $O/sanitize_pytest.log; tail -3 $O/sanitize_pytest.log; tail -12 $O/sanitize_memcheck.log
timeout 300 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 400 $O/bench_c2.err
timeout 600 python bench.py --workload c3 --steps 50 --warmup 10 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 400 $O/bench_c3.err
python - <<'PY'
import json
for w in ("c2","c3"):
    try:
        d=json.load(open("gpurun_out/r02b/bench_%s.json"%w)); print(w, d["ms_per_step"], d["value"], d["roofline"]["per_kernel_ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(w, "failed", e)
PY
