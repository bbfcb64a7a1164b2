#!/usr/bin/env python
"""SASS of the default neighbour-pass kernels, inner loops marked: python scripts/sass_excerpt.py > profiles/r02/sass_default_kernels.txt
For each kernel: every loop (backward branch) with its instruction count and opcode histogram, then the body of the loop
that holds the most packed-math instructions (FFMA2/FADD2/FMUL2), i.e. the candidate / neighbour loop."""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "sph-erosion_b200", "build", "sph.o")
KERNELS = [("k_density_listILb0ELb1ELi56ELi64ELi4E", "k_density_list<REC=0, PF=1, CAP=56, THREADS=64, UNROLL=4> (default density pass): the loop tests 4 candidates for 2 targets"),
           ("k_force_listILb0ELb0ELb0E", "k_force_list<DIAG=0, REC=0, PF=0> (default force pass): the loop handles 1 list entry for 2 targets")]
names = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
funcs = re.findall(r"Function : (\S+)", names)
for key, title in KERNELS:
    fn = [f for f in funcs if key in f][0]
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, OBJ], capture_output=True, text=True).stdout
    ins = []
    for l in sass.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("=" * 110); print(title); print(fn); print("%d SASS instructions" % len(ins))
    loops = []
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= a: loops.append((int(m.group(1), 16), a))
    best = None
    for lo, hi in loops:
        body = [t for a, t in ins if lo <= a <= hi]
        ops = collections.Counter((t.split(None, 1)[1] if t.startswith("@") else t).split()[0] for t in body)
        packed = sum(v for k, v in ops.items() if k in ("FFMA2", "FADD2", "FMUL2"))
        print("loop 0x%04x..0x%04x: %3d instructions, packed fp32 %d | %s" % (lo, hi, len(body), packed, ", ".join("%s %d" % kv for kv in ops.most_common(12))))
        if len(body) <= 150 and (best is None or packed > best[0] or (packed == best[0] and len(body) < best[3])): best = (packed, lo, hi, len(body))
    _, lo, hi, nb = best
    print("-- body of loop 0x%04x..0x%04x (%d instructions):" % (lo, hi, nb))
    for a, t in ins:
        if lo <= a <= hi: print("    /*%04x*/ %s" % (a, t))
