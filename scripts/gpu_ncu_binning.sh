set -x
mkdir -p gpurun_out/ncu_r02
O=gpurun_out/ncu_r02
ARGS="--steps 6 --warmup 3 --no-cpu-baseline --no-parity-gate --no-reference-gravity"
for K in k_hash k_scatter k_scan_onepass; do
timeout 900 ncu --set full --clock-control none -k regex:$K -s 170 -c 1 -f -o $O/prof_${K}_c3 python bench.py $ARGS > $O/ncu_${K}.log 2>&1
tail -1 $O/ncu_${K}.log
done
