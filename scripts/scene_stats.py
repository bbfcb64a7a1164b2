"""Prints how a benchmark scene evolves: KE, max speed, density range, mean neighbour count."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
workload, steps, every = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dv, fv = (int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "0,0").split(","))
pkg = importlib.import_module("sph-erosion_b200")
n_axis, jitter, _, _ = bench.WORKLOADS[workload]
pos, L = bench.scaled_dam_break(n_axis, jitter)
sim = pkg.FluidSystemSPH(); sim.params.len = L; sim.SetDeltaTime(0.01); sim.params.g[1] = bench.scene_gravity(n_axis); sim.set_variant(dv, fv)
sim.upload_state(pos, np.zeros_like(pos))
for step in range(1, steps + 1):
    sim.Run()
    if step % every == 0 or step == 1:
        v = sim.download("vel").astype(np.float64); rho = sim.download("density"); p = sim.download("pos")
        ns, _ = sim.debug_neighbours() if pos.shape[0] <= 1100000 else (None, None)
        sp = np.sqrt((v * v).sum(1))
        print("step %4d  meanKE %.4e  max|v| %.3f  rho [%.0f, %.0f] mean %.0f  nbr/particle %s  y-range [%.3f, %.3f] finite %s" % (
            step, (0.5 * 0.02 * sp * sp).mean(), sp.max(), rho.min(), rho.max(), rho.mean(),
            "%.1f" % (ns[-1] / pos.shape[0]) if ns is not None else "-", p[:, 1].min(), p[:, 1].max(), np.isfinite(p).all()), flush=True)
