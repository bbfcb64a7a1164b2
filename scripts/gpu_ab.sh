# usage: V="3,3 7,7 9,9" W=c2 bash scripts/gpu_ab.sh  -- parity tests of the kernel variants, then an A/B sweep
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8
for w in ${W:-c2 c3-noterrain}; do timeout 600 python scripts/sweep.py $w ${STEPS:-50} "${V:-3,3 7,7 9,9}" 2>&1 | grep "^variant" | tee -a gpurun_out/sweep_$w.log; done
