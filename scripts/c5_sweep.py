"""BASELINE configs[4] (SURVEY.md 8d "C5"): neighbour-search / force microbench sweep.
N in {100^3 .. 400^3} jittered lattice particles at spacing 0.025, smoothing radius h in {0.0482, 0.0607, 0.0765}
(about 30 / 60 / 120 neighbours), DEFAULT kernel variants (6,3) unless V is set; one JSON line per point with per-kernel ms, particle-updates/s, the measured mean
neighbour count, algorithmic HBM GB/s (DESIGN.md bytes per particle) and the secondary "gather" figure
(neighbour-candidate reads x 16 B, served from L1/L2 -- NOT DRAM traffic).  Usage: python scripts/c5_sweep.py [out.jsonl]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sph-erosion_b200")
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
axes = [int(a) for a in os.environ.get("AXES", "100 160 256 320 400").split()]
hs = [float(h) for h in os.environ.get("HS", "0.0482 0.0607 0.0765").split()]
variants = [tuple(int(x) for x in v.split(",")) for v in os.environ.get("V", "6,3").split()]
steps = int(os.environ.get("STEPS", "10"))
peak, _ = bench.measured_peak()
for n_axis in axes:
    pos, L = bench.scaled_dam_break(n_axis, jitter=True)
    n = pos.shape[0]
    for h in hs:
        for dv, fv in variants:
            sim = pkg.FluidSystemSPH()
            sim.params.len = L; sim.params.h = h; sim.params.g[1] = bench.scene_gravity(n_axis); sim.SetDeltaTime(0.0)  # dt = 0: static scene, every step identical
            sim.set_variant(dv, fv)
            if "CAP" in os.environ: sim.set_nlist_capacity(int(os.environ["CAP"]))
            sim.upload_state(pos, np.zeros_like(pos))
            if n <= 40_000_000: sim.set_l2_flush(256 << 20)
            for _ in range(int(os.environ.get("ADAPT", "16"))):      # the list sizing adapts from counters the host reads between calls
                sim.timed_steps(1, per_kernel=False)
            ms, pk, _ = sim.timed_steps(steps)
            nb = sim.debug_neighbours_total() / n if n <= 33_000_000 else None
            gi = sim.grid_info()
            cells = gi.dim[0] * gi.dim[1] * gi.dim[2]
            cand = 27.0 * n / max(cells * (L * 2) ** 3 / ((gi.dim[0] * gi.cell) * (gi.dim[1] * gi.cell) * (gi.dim[2] * gi.cell)), 1)  # ~ candidates per particle per pass in the filled region
            t = {k: v / steps for k, v in pk.items()}
            line = {"particles": n, "h": h, "mean_neighbours": nb, "variant": [dv, fv], "list_capacity": "%d/%d" % (sim.nlist_capacity(), sim.nlist_smem_entries()), "ms_per_step": ms / steps,
                    "particle_updates_per_s": n / (ms / steps * 1e-3), "per_kernel_ms": t,
                    "hbm_GBps_algorithmic": {k: bench.ALGO_BYTES[k] * n / (t[k] * 1e-3) / 1e9 for k in ("hash", "scatter", "reorder", "density", "force") if t[k] > 0},
                    "algorithmic_bytes": "SURVEY.md 8(d): hash 24, scan 6, scatter 52, reorder 68, density 20, force 64",
                    "hbm_frac_step": 234 * n / (ms / steps * 1e-3) / 1e9 / peak}
            print(json.dumps(line), flush=True)
            if out: out.write(json.dumps(line) + "\n"); out.flush()
            del sim
