"""A/B sweep of kernel variants on one scene: per-kernel ms and max relative difference of the
resulting state against variant (0,0).  Usage: python scripts/sweep.py c3 30 "0,0 1,0 4,0 ..." """
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

workload, steps = sys.argv[1], int(sys.argv[2])
variants = [tuple(int(x) for x in v.split(",")) for v in sys.argv[3].split()]
pkg = importlib.import_module("sph-erosion_b200")
n_axis, jitter, _, _ = bench.WORKLOADS[workload]
pos, L = bench.scaled_dam_break(n_axis, jitter)
n = pos.shape[0]
ref = None
for dv, fv in variants:
    sim = pkg.FluidSystemSPH()
    sim.params.len = L
    sim.SetDeltaTime(0.01); sim.params.g[1] = bench.scene_gravity(n_axis)
    sim.set_variant(dv, fv)
    sim.upload_state(pos, np.zeros_like(pos))
    sim.set_l2_flush(256 << 20)
    sim.timed_steps(5, per_kernel=False)
    ms, pk, _ = sim.timed_steps(steps)
    rho = sim.download("density"); p = sim.download("pos")
    if ref is None:
        ref = (rho, p)
    drho = np.abs(rho - ref[0]).max() / np.abs(ref[0]).max()
    dpos = np.abs(p - ref[1]).max() / np.abs(ref[1]).max()
    ns = None
    print("variant d=%d f=%d  ms/step %.4f  density %.4f  force %.4f  binning %.4f | drho %.2e dpos %.2e" % (
        dv, fv, ms / steps, pk["density"] / steps, pk["force"] / steps,
        (pk["hash"] + pk["scan"] + pk["scatter"] + pk["reorder"]) / steps, drho, dpos), flush=True)
    del sim
