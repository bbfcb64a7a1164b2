import importlib, sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
pkg = importlib.import_module("sph-erosion_b200")
n_axis, jitter, terrain, desc = bench.WORKLOADS["c3"]
pos, L = bench.scaled_dam_break(n_axis, jitter)
sim = pkg.FluidSystemSPH(device=0)
sim.params.len = L; sim.params.g[1] = bench.scene_gravity(n_axis); sim.SetDeltaTime(0.01)
sim.upload_state(pos, np.zeros_like(pos))
grid, tinfo = bench.attach_terrain(pkg, L, n_axis)
for k in range(6):
    sim.timed_steps(50, grid=grid, per_kernel=False)
    p = sim.download("pos"); v = sim.download("vel"); r = sim.download("density")
    bad = ~np.isfinite(p).all(1) | ~np.isfinite(v).all(1)
    print("after %d steps: non-finite particles %d, |v| max %.3f, x range %.3f..%.3f, on -x wall %d, rho max %.1f" % (50*(k+1), bad.sum(), np.abs(v[~bad]).max(), p[~bad,0].min(), p[~bad,0].max(), (p[:,0]==-np.float32(L)).sum(), r.max()), flush=True)
    if bad.any():
        i = np.nonzero(bad)[0][:5]; print(i, p[i], v[i])
