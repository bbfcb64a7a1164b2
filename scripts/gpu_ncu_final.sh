# final evidence pack (final kernels of round 2): launch list of the default bench command + one --set full capture per kernel
set -x
mkdir -p gpurun_out/ncu_final
O=gpurun_out/ncu_final
ARGS="--steps 6 --warmup 3 --no-cpu-baseline --no-parity-gate --no-reference-gravity"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c3_all.csv python bench.py $ARGS > $O/launches_c3.log 2>&1
tail -1 $O/launches_c3.log
for K in k_density_list k_force_list k_terrain_contact k_rank_reorder k_scatter k_scan_onepass k_hash k_terrain_grant k_terrain_apply; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 155 -c 1 -f -o $O/prof_${K}_c3 python bench.py $ARGS > $O/ncu_${K}.log 2>&1
tail -1 $O/ncu_${K}.log
done
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
ls -la $O
