"""Multi-PROCESS check of the peer-memory slab exchange (run under torchrun, one rank per GPU; with
SPHE_ONE_GPU=1 all ranks share cuda:0, which exercises the same CUDA-IPC mapping on a single-GPU box).

Every rank steps its x-slab with PeerSlabDriver (mailbox stores over NVLink / IPC, device-side flag waits,
terrain accumulators summed with NCCL or gloo all-reduce); rank 0 also runs the whole scene on one handle.
Bar: positions, velocities, densities, carried sediment and terrain heights BIT-EQUAL to the single-handle
run, sum(heights) + sum(sediment) exactly conserved."""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
one_gpu = os.environ.get("SPHE_ONE_GPU") == "1"
devno = 0 if one_gpu else local
torch.cuda.set_device(devno)
dev = torch.device("cuda", devno)
if one_gpu:
    dist.init_process_group("gloo")      # NCCL refuses two ranks on one device; the exchange itself does not use it
    def reduce(t):
        h = t.cpu(); dist.all_reduce(h); t.copy_(h)
else:
    dist.init_process_group("nccl", device_id=dev)
    def reduce(t):
        dist.all_reduce(t)
pkg = importlib.import_module("sph-erosion_b200"); slabs = importlib.import_module("sph-erosion_b200.slabs")
import test_gpu_slabs as T

steps = int(os.environ.get("STEPS", "12"))
box = (1.2, 0.3, 0.3)
params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
grid, pos, vel = T._terrain_scene(pkg)
n = pos.shape[0]
cap = 1 << 15
sim, backend, cols = slabs.make_gpu_slab(pkg, devno, rank, world, box, params, None, cap, (6, 3))
order = np.argsort(pos[:, 0], kind="stable")
part = np.array_split(order, world)[rank]
sim.slab_upload(pos[part], vel[part], part.astype(np.int32))
if one_gpu:
    sim.slab_peer_timeout(400_000_000_000)   # ranks time-slice one GPU: a waiting kernel can sit out whole time slices
total0 = grid.total_fx()
mode = os.environ.get("TERRAIN_SHARE", "window")   # the windowed zone sums travel through the mailboxes: no NCCL P2P needed, also on one GPU
cell_t = 2.4 / 256
if mode == "window" and 256 // world < 2 * slabs.terrain_margin_rows(sim.grid_info().cell, cell_t):
    mode = "allreduce"      # slabs narrower than two boundary zones: the windowed scheme does not apply
if mode == "window":
    share = slabs.TerrainWindowShare(grid, dev, rank, world, slabs.terrain_row_cuts(sim.grid_info(), cols, -1.2, cell_t),
                                     slabs.terrain_margin_rows(sim.grid_info().cell, cell_t), dist=dist, peer=True)
else:
    share = slabs.TerrainShare(grid, dev, reduce)
drv = slabs.PeerSlabDriver(sim, rank, world, cap, n + 4 * cap, share)
drv.connect(dist)
for _ in range(steps):
    drv.step()
info = drv.drain()
ids, p, v, rho, sed = sim.slab_download()
sed_fx = sim.sediment_total_fx()
gathered = [None] * world
own = share.own if mode == "window" else ((0, grid.shape()[0]) if rank == 0 else (0, 0))   # replicas: count the terrain once
win = share.window if mode == "window" else (0, grid.shape()[0])
dist.all_gather_object(gathered, (ids, p, v, rho, sed_fx, grid.heights_fx(), grid.contacts(), info, own, win, grid.window_violations()))
ok = True
if rank == 0:
    one = T._single(pkg, box, params, (6, 3), pos, vel)
    g1, _, _ = T._terrain_scene(pkg)
    for _ in range(steps):
        one.Run(g1)
    allids = np.concatenate([g[0] for g in gathered]); o = np.argsort(allids)
    assert np.array_equal(allids[o], np.arange(n)), "every particle owned exactly once"
    for j, name in ((1, "pos"), (2, "vel"), (3, "density")):
        a = np.concatenate([g[j] for g in gathered])[o]
        b = one.download(name)
        same = np.array_equal(a, b)
        ok &= same
        print("%-8s bit-equal: %s" % (name, same))
    h1 = g1.heights_fx()
    for r, g in enumerate(gathered):
        w0, w1 = g[9]
        same = np.array_equal(g[5][w0:w1], h1[w0:w1]) and g[10] == 0
        ok &= same
        print("terrain rows [%d,%d) kept by rank %d bit-equal to the single-GPU terrain, no window violations: %s" % (w0, w1, r, same))
    sed_k = sum(g[4] for g in gathered)
    cons = sum(int(g[5][g[8][0]:g[8][1]].astype(np.int64).sum()) for g in gathered) + sed_k == total0
    ok &= bool(cons) and sed_k == one.sediment_total_fx() and sed_k > 0
    print("sediment in flight %d (single GPU %d), conservation exact: %s" % (sed_k, one.sediment_total_fx(), cons))
    contacts = sum(g[6] for g in gathered)
    ok &= contacts == g1.contacts() and contacts > 1000
    print("contacts %d (single GPU %d); exchange counts per rank: %s" % (contacts, g1.contacts(), [g[7] for g in gathered]))
    print("PEER_CHECK %s world=%d one_gpu=%s terrain_share=%s" % ("OK" if ok else "FAILED", world, one_gpu, mode))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
