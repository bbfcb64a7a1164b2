# launch list of the default bench command (per-launch device time, cold-cache and serialised: compare SHARES)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1900} -c ${COUNT:-110} --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity-gate --no-reference-gravity ${ARGS} > gpurun_out/launches_${TAG}.log 2>&1
tail -3 gpurun_out/launches_${TAG}.log
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_${TAG}.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
t=collections.defaultdict(list)
for r in rows[1:]:
    try: t[r[ik].split("(")[0][:60]].append(float(r[iv].replace(",","")))
    except ValueError: pass
tot=sum(sum(v) for v in t.values())
for k,v in sorted(t.items(), key=lambda kv:-sum(kv[1])): print("%-62s n=%3d mean=%9.1f ns share=%5.1f%%" % (k,len(v),sum(v)/len(v),100*sum(v)/tot))
PY
