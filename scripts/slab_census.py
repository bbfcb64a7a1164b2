"""Particle census of the slab exchange on the weak-scaling bench scene (slabs.channel_block): every
`EVERY` steps all owned ids are gathered and checked -- each particle owned exactly once.  Lost or
duplicated ids are reported with their last known position / owner.

  one process, K slabs on one GPU (LocalPeerGroup):   K=8 NAXIS=40 STEPS=120 python scripts/slab_census.py
  one rank per GPU (PeerSlabDriver):                  torchrun --nproc-per-node 8 scripts/slab_census.py
"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench  # noqa: E402
pkg = importlib.import_module("sph-erosion_b200"); slabs = importlib.import_module("sph-erosion_b200.slabs")

n_axis = int(os.environ.get("NAXIS", "40")); steps = int(os.environ.get("STEPS", "120")); every = int(os.environ.get("EVERY", "10"))
multi = "RANK" in os.environ
gy = bench.scene_gravity(n_axis)
layer = int(n_axis * n_axis * (0.0457 * 1.001 / 0.025 + 1))
cap = max(2 * slabs.HALO * layer, 1 << 14)


def census(step, ids_by_rank, pos_by_rank, prev):
    ids = np.concatenate(ids_by_rank)
    owner = np.concatenate([np.full(len(x), r) for r, x in enumerate(ids_by_rank)])
    pos = np.concatenate(pos_by_rank)
    total = prev["n"]
    cnt = np.bincount(ids, minlength=total)
    lost = np.nonzero(cnt == 0)[0]; dup = np.nonzero(cnt > 1)[0]
    print("step %4d: owned %d of %d, missing %d (in transit %d), duplicated %d, per rank %s" % (step, len(ids), total, len(lost), prev.get("transit", 0), len(dup), [len(x) for x in ids_by_rank]), flush=True)
    for name, arr in (("lost", lost), ("dup", dup)):
        for i in arr[:8]:
            print("   %s id %d: previous owner %d pos %s" % (name, i, prev["owner"][i], prev["pos"][i]), flush=True)
    o = np.argsort(ids, kind="stable")
    u, first = np.unique(ids[o], return_index=True)
    prev["owner"][u] = owner[o][first]; prev["pos"][u] = pos[o][first]
    return abs(len(lost) - prev.get("transit", 0)) + len(dup)


jitter = os.environ.get("JITTER", "0") == "1"; terrain = os.environ.get("TERRAIN", "0") == "1"
if not multi:
    K = int(os.environ.get("K", "8"))
    sims = []; n_tot = 0; allpos = []; shares = None
    for r in range(K):
        pos, ids, box, bounds = slabs.channel_block(n_axis, K, r, jitter, layout=os.environ.get('LAYOUT', 'contiguous'))
        sim, b, cols = slabs.make_gpu_slab(pkg, 0, r, K, box, dict(len=box[1], dt=0.01, g=(0.0, gy, 0.0)), bounds, cap)
        sim.slab_upload(pos, np.zeros_like(pos), ids); sims.append(sim); n_tot += len(ids); allpos.append(pos)
        print("rank", r, "cols", cols[r])
    if terrain:
        shares = []
        for r in range(K):
            grid, tinfo = bench.attach_terrain(pkg, box[1], n_axis, nx_mult=K)
            gi = sims[0].grid_info()
            shares.append(slabs.TerrainWindowShare(grid, torch.device("cuda", 0), r, K, slabs.terrain_row_cuts(gi, cols, tinfo["terrain_origin"][0], tinfo["terrain_cell"]),
                                                   slabs.terrain_margin_rows(gi.cell, tinfo["terrain_cell"]), swap=False))
    group = slabs.LocalPeerGroup(sims, cap, int(n_axis ** 3 * 1.3) + 6 * cap, shares=shares)
    prev = {"n": n_tot, "owner": np.repeat(np.arange(K), n_axis ** 3), "pos": np.concatenate(allpos)}
    bad = 0
    for s in range(1, steps + 1):
        group.step()
        if s % every == 0:
            group.drain()
            got = [x.slab_download() for x in sims]
            tr = [x.slab_transit() for x in sims]
            prev["transit"] = sum(t["to_left"] + t["to_right"] for t in tr); prev["forwarded"] = sum(t["forwarded"] for t in tr)
            bad += census(s, [g[0] for g in got], [g[1] for g in got], prev)
            vmax = max(float(np.abs(g[2]).max()) for g in got)
            print("          max |v| component %.2f -> %.1f neighbour-grid columns per step%s" % (
                vmax, vmax * 0.01 / 0.0457, ", window violations %s" % [t.grid.window_violations() for t in shares] if shares else ""), flush=True)
    print("CENSUS", "OK" if bad == 0 else "FAILED", "forwarded records:", prev.get("forwarded"))
else:
    import torch.distributed as dist
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pos, ids, box, bounds = slabs.channel_block(n_axis, world, rank, False, layout=os.environ.get('LAYOUT', 'contiguous'))
    sim, b, cols = slabs.make_gpu_slab(pkg, local, rank, world, box, dict(len=box[1], dt=0.01, g=(0.0, gy, 0.0)), bounds, cap)
    sim.slab_upload(pos, np.zeros_like(pos), ids)
    drv = slabs.PeerSlabDriver(sim, rank, world, cap, int(n_axis ** 3 * 1.3) + 6 * cap)
    drv.connect(dist)
    prev = None
    if rank == 0:
        allpos = np.concatenate([slabs.channel_block(n_axis, world, r, False, layout=os.environ.get('LAYOUT', 'contiguous'))[0] for r in range(world)])
        prev = {"n": world * n_axis ** 3, "owner": np.repeat(np.arange(world), n_axis ** 3), "pos": allpos}
    bad = 0
    for s in range(1, steps + 1):
        drv.step()
        if s % every == 0:
            drv.drain()
            g = sim.slab_download()
            out = [None] * world
            t = sim.slab_transit()
            dist.all_gather_object(out, (g[0], g[1], t["to_left"] + t["to_right"]))
            if rank == 0:
                prev["transit"] = sum(x[2] for x in out)
                bad += census(s, [x[0] for x in out], [x[1] for x in out], prev)
    if rank == 0:
        print("CENSUS", "OK" if bad == 0 else "FAILED")
    dist.barrier(); dist.destroy_process_group()
