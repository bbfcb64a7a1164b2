# A/B of build flags on the C5 high-neighbour points: CONFIGS="flags1|flags2"
mkdir -p gpurun_out/abc5
IFS='|' read -ra CFG <<< "$CONFIGS"
i=0
for F in "${CFG[@]}"; do
  SPHE_NVCC_EXTRA="$F" python sph-erosion_b200/build.py > gpurun_out/abc5/build_$i.log 2>&1 || tail -5 gpurun_out/abc5/build_$i.log
  echo "== [$F]"
  AXES="${AXES:-100 160}" HS="${HS:-0.0765 0.0607}" timeout 900 python scripts/c5_sweep.py gpurun_out/abc5/sweep_$i.jsonl 2> gpurun_out/abc5/sweep_$i.err | python scripts/c5_fmt.py
  i=$((i+1))
done
python sph-erosion_b200/build.py > /dev/null 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -3
