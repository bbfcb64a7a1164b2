set -x
mkdir -p gpurun_out
W=${W:-c2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 60 -c 1 -f -o gpurun_out/prof_force_$W python bench.py --workload $W --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_force.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_density -s 60 -c 1 -f -o gpurun_out/prof_density_$W python bench.py --workload $W --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_density.log 2>&1
ls -la gpurun_out
