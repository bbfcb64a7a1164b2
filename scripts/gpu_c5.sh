# C5 (BASELINE configs[4]): 1M .. 64M particles x 3 smoothing radii, default kernel variants; then DRAM bytes per launch from ncu for a subset
set -x
mkdir -p gpurun_out/c5
timeout 1500 python scripts/c5_sweep.py gpurun_out/c5/c5_sweep.jsonl 2> gpurun_out/c5/c5_sweep.err | python scripts/c5_fmt.py
tail -3 gpurun_out/c5/c5_sweep.err
AXES="100 256" STEPS=2 ADAPT=8 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_density_list|k_force_list' --csv --log-file gpurun_out/c5/c5_ncu_dram.csv python scripts/c5_sweep.py gpurun_out/c5/c5_sweep_under_ncu.jsonl > /dev/null 2> gpurun_out/c5/c5_ncu.err
tail -2 gpurun_out/c5/c5_ncu.err; wc -l gpurun_out/c5/c5_ncu_dram.csv
