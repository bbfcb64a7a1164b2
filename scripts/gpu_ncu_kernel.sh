# usage: K=<kernel regex> W=<workload> DV=<density variant> FV=<force variant> SKIP=<launches to skip> TAG=<name>
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-30} -c 1 -f -o gpurun_out/prof_$TAG python bench.py --workload ${W:-c3} --steps ${SKIP:-30} --warmup 5 --no-cpu-baseline --density-variant ${DV:-0} --force-variant ${FV:-0} > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
