"""Per-kernel times, mean neighbour count and list overflow along the dam-break transient of a bench workload.
Usage: python scripts/transient.py c2 [chunks=6] [steps_per_chunk=50] [density_variant force_variant]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sph-erosion_b200")
w = sys.argv[1]; chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 6; per = int(sys.argv[3]) if len(sys.argv) > 3 else 50
n_axis, jitter, terrain, _ = bench.WORKLOADS[w]
pos, L = bench.scaled_dam_break(n_axis, jitter)
n = pos.shape[0]
sim = pkg.FluidSystemSPH()
sim.params.len = L; sim.params.g[1] = bench.scene_gravity(n_axis); sim.SetDeltaTime(0.01)
if len(sys.argv) > 5: sim.set_variant(int(sys.argv[4]), int(sys.argv[5]))
sim.upload_state(pos, np.zeros_like(pos))
sim.set_l2_flush(256 << 20)
grid = bench.attach_terrain(pkg, L, n_axis)[0] if terrain else None
for c in range(chunks):
    ms, pk, _ = sim.timed_steps(per, grid=grid)
    nb = sim.debug_neighbours_total() / n
    rho = sim.download("density")
    print("steps %4d-%4d: ms/step %.4f density %.4f force %.4f terrain %.4f binning %.4f | mean nbrs %.1f rho max %.0f mean %.0f | overflowed pairs %d (%.2f%%) cap %s" % (
        c * per, (c + 1) * per, ms / per, pk["density"] / per, pk["force"] / per, pk["terrain"] / per,
        (pk["hash"] + pk["scan"] + pk["scatter"] + pk["reorder"]) / per, nb, rho.max(), rho.mean(), sim.nlist_overflowed(), 200.0 * sim.nlist_overflowed() / n, "%d/%d" % (sim.nlist_capacity(), sim.nlist_smem_entries())), flush=True)
