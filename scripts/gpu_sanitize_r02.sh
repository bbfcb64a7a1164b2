# compute-sanitizer over the kernels written or changed in round 2 (staged passes, in-place exchange, zone sums, one-pass
# scan, scatter rounds, class-sorted terrain stage): memcheck over a wide subset, racecheck over the shared-memory users
set -x
mkdir -p gpurun_out/san
O=gpurun_out/san
SEL='dense_neighbourhoods or golden_step1 or (one_cell and (default or staged)) or odd_counts or (k_slabs and 3 and not tpp) or (sharing_one_eroding and 2) or (lockstep_vs_oracle and (default or staged)) or pair_masks and lattice or device_buffer'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file $O/sanitize_memcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_slabs.py tests/test_gpu_terrain.py -x -q -k "$SEL" > $O/sanitize_memcheck_pytest.log 2>&1
echo "rc=$?" >> $O/sanitize_memcheck_pytest.log; tail -3 $O/sanitize_memcheck_pytest.log; tail -4 $O/sanitize_memcheck.log
SEL2='(golden_step1 and (default or staged)) or (one_cell and staged) or (k_slabs and 3 and default and async) or (lockstep_vs_oracle and default)'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 --log-file $O/sanitize_racecheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_slabs.py tests/test_gpu_terrain.py -x -q -k "$SEL2" > $O/sanitize_racecheck_pytest.log 2>&1
echo "rc=$?" >> $O/sanitize_racecheck_pytest.log; tail -3 $O/sanitize_racecheck_pytest.log; tail -4 $O/sanitize_racecheck.log
