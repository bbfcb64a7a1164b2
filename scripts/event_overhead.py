import importlib, sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
pkg = importlib.import_module("sph-erosion_b200")
for w in ("c2",):
    n_axis, jitter, terrain, _ = bench.WORKLOADS[w]
    pos, L = bench.scaled_dam_break(n_axis, jitter)
    for pk in (True, False, True, False):
        sim = pkg.FluidSystemSPH(); sim.params.len = L; sim.params.g[1] = bench.scene_gravity(n_axis); sim.SetDeltaTime(0.01)
        sim.upload_state(pos, np.zeros_like(pos)); sim.set_l2_flush(256 << 20)
        sim.timed_steps(10, per_kernel=False)
        ms, k, _ = sim.timed_steps(100, per_kernel=pk)
        print(w, "per_kernel", pk, "ms/step %.4f" % (ms / 100), "sum of kernels %.4f" % (sum(k.values()) / 100))
