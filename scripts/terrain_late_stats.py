"""The terrain stage once the c3 fluid has spread (step 600): survivors of the cull per Grid::collision class, contacts per
step, and how the survivors split by vertical velocity and by height above their own cell's four vertices."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sph-erosion_b200")
n_axis, jitter, terrain, _ = bench.WORKLOADS["c3"]
pos, L = bench.scaled_dam_break(n_axis, jitter)
n = pos.shape[0]
sim = pkg.FluidSystemSPH()
sim.params.len = L; sim.params.g[1] = bench.scene_gravity(n_axis); sim.SetDeltaTime(0.01)
sim.upload_state(pos, np.zeros_like(pos))
grid, tinfo = bench.attach_terrain(pkg, L, n_axis)
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 600
for _ in range(settle): sim.Run(grid)
p0 = sim.download("pos"); v0 = sim.download("vel")
c0 = grid.contacts(reset=True)
ms, pk, _ = sim.timed_steps(20, grid=grid)
print("step %d: terrain stage %.4f ms/step, survivors per class %s, contacts/step %.0f" % (settle, pk["terrain"] / 20, sim.terrain_survivors(), grid.contacts() / 20))
# geometry of the state at `settle` (host side, approximate: one step later positions ~ p0 + v0 dt)
h = grid.heights(); o = np.array(tinfo["terrain_origin"], np.float32); sc = np.float32(tinfo["terrain_cell"])
pn = p0 + v0 * np.float32(0.01)
tx = (pn[:, 0] - o[0]) / sc; tz = (pn[:, 2] - o[2]) / sc; ty = (pn[:, 1] - o[1]) / sc
ix = np.clip(np.floor(tx).astype(int), 0, h.shape[0] - 2); iz = np.clip(np.floor(tz).astype(int), 0, h.shape[1] - 2)
m4 = np.maximum.reduce([h[ix, iz], h[ix + 1, iz], h[ix, iz + 1], h[ix + 1, iz + 1]])
pad = np.pad(h, 2, mode="edge")
m16 = np.maximum.reduce([pad[ix + 2 + dx, iz + 2 + dz] for dx in (-1, 0, 1, 2) for dz in (-1, 0, 1, 2)])
surv = ty <= m16 + 0.01
up = v0[:, 1] > 0
above4 = ty > m4 + 0.01
print("particles %d, below 4x4 max (cull survivors, approx) %d" % (n, surv.sum()))
print("  of those: above own cell's 4 vertices %d (moving up %d, not up %d); not above %d" % ((surv & above4).sum(), (surv & above4 & up).sum(), (surv & above4 & ~up).sum(), (surv & ~above4).sum()))
d = np.abs(v0[:, [0, 2]] * 0.01 / sc)
print("  displacement per step in terrain cells: median %.2f, 90%% %.2f" % (np.median(np.linalg.norm(d, axis=1)[surv]), np.percentile(np.linalg.norm(d, axis=1)[surv], 90)))
print("  height above the own-cell max, survivors: quartiles", np.percentile((ty - m4)[surv], [5, 25, 50, 75, 95]).round(2))
