# final evidence pack of round 2: launch list of the default bench command + one --set full capture per dominant kernel
set -x
mkdir -p gpurun_out/ncu_r02
O=gpurun_out/ncu_r02
ARGS="--steps 6 --warmup 3 --no-cpu-baseline --no-parity-gate --no-reference-gravity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1770 -c 66 --csv --log-file $O/launches_c3.csv python bench.py $ARGS > $O/launches_c3.log 2>&1
for K in k_density_list k_force_list k_terrain_contact k_rank_reorder k_slab_classify; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 170 -c 1 -f -o $O/prof_${K}_c3 python bench.py $ARGS > $O/ncu_${K}.log 2>&1
tail -1 $O/ncu_${K}.log
done
ls -la $O
