set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_terrain_contact -s 160 -c 1 -f -o gpurun_out/prof_terrain_contact_c3 python bench.py --workload c3 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_terrain.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 60 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
tail -3 gpurun_out/ncu_terrain.log
