#!/usr/bin/env python
"""Hot spots of an ncu source page (SASS view): python scripts/ncu_hot2.py prof.ncu-rep [top]
Prints the instructions with the most stall samples and the most executed instructions, plus region totals."""
import csv, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in data); tot_i = sum(f(r, "Instructions Executed") for r in data)
print("total samples %d, total warp instructions %d, rows %d" % (tot_s, tot_i, len(data)))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("-- by address (all): idx samples% inst% avgthr  top-stall  sass")
for i, r in enumerate(data):
    s = f(r, "# Samples"); n = f(r, "Instructions Executed")
    if s / max(tot_s, 1) > 0.004 or n / max(tot_i, 1) > 0.004 or "BRA" in r[ix["Source"]] or "BAR" in r[ix["Source"]]:
        st = max(stall_cols, key=lambda c: f(r, c))
        print("%4d %5.2f %5.2f %5.1f %-18s %s" % (i, 100 * s / max(tot_s, 1), 100 * n / max(tot_i, 1), f(r, "Avg. Threads Executed"), st if f(r, st) > 0 else "", r[ix["Source"]].strip()[:90]))
