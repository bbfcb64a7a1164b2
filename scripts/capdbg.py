import importlib, sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
pkg = importlib.import_module("sph-erosion_b200")
pos, L = bench.scaled_dam_break(100, jitter=True)
for h in (0.0607, 0.0765):
    sim = pkg.FluidSystemSPH(); sim.params.len = L; sim.params.h = h; sim.SetDeltaTime(0.0); sim.set_variant(3, 3)
    sim.upload_state(pos, np.zeros_like(pos))
    tr = []
    for k in range(30):
        ms, pk, _ = sim.timed_steps(1)
        tr.append("%d/%d:%.2f+%.2f" % (sim.nlist_capacity(), sim.nlist_smem_entries(), pk["density"], pk["force"]))
    print(h, " ".join(tr))
