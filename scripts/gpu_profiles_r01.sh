# Round-1 profile set: launch lists (c2, c3), ncu --set full of the dominant kernels, C5 sweep.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1850 -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 3 --warmup 3 --settle 150 --no-cpu-baseline > gpurun_out/ncu_bench_c3.log 2>&1
for K in k_density_list k_force_list k_rank_reorder; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 1 -f -o gpurun_out/prof_${K}_c2 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_$K.log 2>&1
done
for K in k_terrain_contact k_density_list k_force_list; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 170 -c 1 -f -o gpurun_out/prof_${K}_c3 python bench.py --workload c3 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_${K}_c3.log 2>&1
done
if [ -n "$SWEEP" ]; then AXES="100 160 256 320" STEPS=40 timeout 900 python scripts/c5_sweep.py gpurun_out/c5_sweep.jsonl 2>&1 | python scripts/c5_fmt.py; fi
ls -la gpurun_out | head -40
