# ncu --set full of the two default (neighbour-list) kernels on the c2 workload
set -x
mkdir -p gpurun_out
W=${W:-c2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force_list -s 20 -c 1 -f -o gpurun_out/prof_force_list_$W python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_force_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_density_list -s 20 -c 1 -f -o gpurun_out/prof_density_list_$W python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_density_list.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_rank_reorder -s 20 -c 1 -f -o gpurun_out/prof_reorder_$W python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_reorder.log 2>&1
ls -la gpurun_out
