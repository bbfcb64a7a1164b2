set -x
mkdir -p gpurun_out/c5
AXES=400 HS="0.0607 0.0765" timeout 1500 python scripts/c5_sweep.py gpurun_out/c5/c5_sweep_64m.jsonl 2> gpurun_out/c5/c5_sweep_64m.err | python scripts/c5_fmt.py
tail -3 gpurun_out/c5/c5_sweep_64m.err
