# usage: bash scripts/gpu_quick.sh "<pytest args>"
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest $1 -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_quick.log
