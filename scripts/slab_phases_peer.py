"""Per-phase device and host times of the multi-GPU step with the peer-memory exchange and the slab-local terrain
(run under torchrun): send (classify + forward + headers), recv (append), phase 0 (binning .. forces .. contact),
zone sum of `want`, phase 1 (grants), zone sum of `delta`, phase 2 (apply + cull map)."""
import importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local); dist.init_process_group("nccl", device_id=dev)
pkg = importlib.import_module("sph-erosion_b200"); slabs = importlib.import_module("sph-erosion_b200.slabs")
W = os.environ.get("W", "c3")
n_axis, jitter, terrain, _ = bench.WORKLOADS[W]
layout = "tiled" if terrain else "contiguous"
pos, ids, box, bounds = slabs.channel_block(n_axis, world, rank, jitter, layout=layout)
gy = bench.scene_gravity(n_axis)
layer = int(n_axis * n_axis * (0.0457 * 1.001 / 0.025 + 1)); cap = max(2 * slabs.HALO * layer, 1 << 14)
sim, b, cols = slabs.make_gpu_slab(pkg, local, rank, world, box, dict(len=box[1], dt=0.01, g=(0.0, gy, 0.0)), bounds, cap)
sim.slab_upload(pos, np.zeros_like(pos), ids)
share = None
if terrain:
    grid, tinfo = bench.attach_terrain(pkg, box[1], n_axis, nx_mult=world)
    gi = sim.grid_info()
    share = slabs.TerrainWindowShare(grid, dev, rank, world, slabs.terrain_row_cuts(gi, cols, tinfo["terrain_origin"][0], tinfo["terrain_cell"]),
                                     slabs.terrain_margin_rows(gi.cell, tinfo["terrain_cell"]), dist=dist, peer=os.environ.get("ZONES", "peer") == "peer")
drv = slabs.PeerSlabDriver(sim, rank, world, cap, int(pos.shape[0] * 1.3) + 6 * cap, share)
drv.connect(dist)
for _ in range(int(os.environ.get("SETTLE", "150")) if terrain else 10): drv.step()
drv.drain()
K = 50
names = ["send", "recv", "phase0", "sum want", "phase1", "sum delta", "phase2"] if terrain else ["send", "recv", "step"]
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(K)]
host = np.zeros((K, len(names)))
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(K):
    e = ev[k]; h = [time.perf_counter()]
    e[0].record()
    sim.slab_send(); e[1].record(); h.append(time.perf_counter())
    drv.tickets.append(sim.slab_recv()); drv.tickets = drv.tickets[-4:]; e[2].record(); h.append(time.perf_counter())
    if terrain:
        sim.step_phase(grid, 0); e[3].record(); h.append(time.perf_counter())
        share.sum_zones(sim, 0); e[4].record(); h.append(time.perf_counter())
        sim.step_phase(grid, 1); e[5].record(); h.append(time.perf_counter())
        share.sum_zones(sim, 1); e[6].record(); h.append(time.perf_counter())
        sim.step_phase(grid, 2); e[7].record(); h.append(time.perf_counter())
    else:
        sim.Run(); e[3].record(); h.append(time.perf_counter())
    host[k] = np.diff(h)
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / K
drv.drain()
devt = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(len(names))] for e in ev])
gap = np.array([ev[k][len(names)].elapsed_time(ev[k + 1][0]) for k in range(K - 1)])
print("rank %d %s wall/step %.3f ms | device ms: %s gap %.3f | host ms: %s" % (
    rank, W, wall * 1e3, " ".join("%s %.3f" % (n, t) for n, t in zip(names, devt.mean(0))), gap.mean(),
    " ".join("%s %.3f" % (n, t) for n, t in zip(names, host.mean(0) * 1e3))), flush=True)
dist.barrier(); dist.destroy_process_group()
