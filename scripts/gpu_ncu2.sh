# usage: W=<workload> TAG=<name> [SKIP=..]; one --set full capture of each staged kernel
set -x
mkdir -p gpurun_out
for K in k_density_stage k_force_stage; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-20} -c 1 -f -o gpurun_out/prof_${K}_$TAG python bench.py --workload ${W:-c2} --steps 22 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1
tail -2 gpurun_out/ncu_${K}_$TAG.log
done
