# compute-sanitizer passes over the kernels (next round: run once per kernel change; ~10x slower than a plain run).
#   bash scripts/gpu_sanitize.sh [memcheck|racecheck|initcheck|synccheck]
set -x
mkdir -p gpurun_out
TOOL=${1:-memcheck}
# small scenes: the default-scene golden test, the one-cell pile-up (spill + row growth + fallback), 3 slabs with ring closure
timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 1 --log-file gpurun_out/sanitize_$TOOL.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_slabs.py -x -q \
    -k "default and (golden_step1 or step1) or one_cell and default or peer_mailbox and 3" 2>&1 | tail -5
tail -20 gpurun_out/sanitize_$TOOL.log
