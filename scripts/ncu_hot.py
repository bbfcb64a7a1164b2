#!/usr/bin/env python
"""Per-SASS-instruction executed counts from an .ncu-rep source page: prints the instruction stream
with executed count (warp-level), avg active threads and stall samples, for regions above a threshold."""
import csv, subprocess, sys
path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {k: hdr.index(k) for k in ["Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"]}
data = rows[hi + 1:]
tot = sum(float(r[ci["Instructions Executed"]] or 0) for r in data)
samp = sum(float(r[ci["# Samples"]] or 0) for r in data)
print("total warp instr %.0f  samples %.0f" % (tot, samp))
for i, r in enumerate(data):
    ie = float(r[ci["Instructions Executed"]] or 0)
    if ie / tot >= thr:
        print("%4d %6.2f%% thr=%5s smp=%5.2f%%  %s" % (i, 100 * ie / tot, r[ci["Avg. Threads Executed"]], 100 * float(r[ci["# Samples"]] or 0) / max(samp, 1), r[ci["Source"]].strip()[:90]))
