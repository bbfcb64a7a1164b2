"""The multi-GPU exchange protocol (sph-erosion_b200/slabs.py SlabDriver + TorchComm) under gloo with
world_size 2 and 3 on CPU.  The CUDA side of each slab is replaced by tests/slab_double.py (numpy +
the oracle); what is under test is the host logic: partitioning, the count swap, the sized record
swap, and that a 2-layer halo gives every owned particle its complete neighbourhood -- the k-slab
run must reproduce the single-domain oracle run BIT FOR BIT, migrations included."""
import multiprocessing as mp
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT, product


def scene(n=2000, seed=5):
    rng = np.random.default_rng(seed)
    pos = np.empty((n, 3), np.float32)
    pos[:, 0] = rng.uniform(-0.55, 0.55, n); pos[:, 1] = rng.uniform(-0.14, 0.0, n); pos[:, 2] = rng.uniform(-0.14, 0.14, n)
    vel = rng.normal(0, 1.5, (n, 3)).astype(np.float32)   # fast enough to cross slab boundaries
    # isolated particles EXACTLY on the -x wall with no x velocity: the reference's box response clamps them to the
    # +x wall (collisionS, fluid_system.h:375-382) -- from the first slab straight to the last one (ring closure)
    wall = np.array([[-0.6, 0.3 + 0.1 * k, -0.4 + 0.2 * k] for k in range(3)], np.float32)
    pos = np.concatenate([pos, wall]); vel = np.concatenate([vel, np.zeros_like(wall)])
    return pos, vel


def squeeze_scene(pos, factor):
    """Most of the particles into the left part of the box: an unbalanced load for equal-width slabs."""
    p = pos.copy()
    p[:, 0] = (-0.55 + (p[:, 0] + 0.55) * np.float32(factor)).astype(np.float32)
    return p


def setup(world):
    from oracle import port
    slabs = __import__("importlib").import_module("sph-erosion_b200.slabs")
    P = port.default_params(dt=0.004, len=0.6)
    G = port.grid_for_box(P, [-0.7] * 3, [0.7] * 3)
    cols = slabs.partition_columns(int(G.dim[0]), world)
    return port, slabs, P, G, cols


def worker(rank, world, port_no, steps, outdir, cap=4096, floor=None, lag=0, extra=None, rebalance_at=None, squeeze=None):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from slab_double import NumpySlabBackend, cell_x
    torch.set_num_threads(1)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port_no, rank=rank, world_size=world)
    port, slabs, P, G, cols = setup(world)
    if floor is not None:
        _ns = slabs.next_size
        slabs.next_size = lambda count, cap_, floor_=floor: max(floor, count // 2)  # too small on purpose -> every step re-sends
    pos, vel = scene()
    if extra is not None:
        pos = np.concatenate([pos, extra[0]]); vel = np.concatenate([vel, extra[1]])
    if squeeze is not None:
        pos = squeeze_scene(pos, squeeze)
    x0, x1 = cols[rank]
    left, right, wrap_l, wrap_r = slabs.ring_links(rank, world)
    b = NumpySlabBackend(P, G, x0, x1, left is not None, right is not None, cap=cap, wrap_left=wrap_l, wrap_right=wrap_r,
                         far_x0=cols[-1][0])
    # deliberately start from a WRONG distribution: particles in the two boundary columns of their
    # owner start on the neighbour across that boundary, as if they had just migrated; the first
    # exchange must send them home and mirror them back as ghosts.
    cx = cell_x(pos[:, 0], G.gmin[0], G.cell, int(G.dim[0]))
    owner = np.searchsorted([c[1] for c in cols], cx, side="right")
    ids = np.arange(len(pos))
    x0s = np.array([c[0] for c in cols]); x1s = np.array([c[1] for c in cols])
    place = owner.copy()
    go_r = (cx >= x1s[owner] - 2) & (owner + 1 < world) & (ids % 3 == 0)
    go_l = (cx < x0s[owner] + 2) & (owner > 0) & (ids % 3 == 1)
    place[go_r] += 1; place[go_l] -= 1
    mine = place == rank
    b.upload(pos[mine], vel[mine], ids[mine])
    drv = slabs.SlabDriver(b, slabs.TorchComm(rank, world), lag=lag, sync_steps=1)
    log = []
    counts = []
    for k in range(steps):
        if rebalance_at is not None and k in rebalance_at:
            drv.drain()

            def reduce(a):
                t = torch.from_numpy(a); dist.all_reduce(t)

            counts.append(b.n_owned)
            cols = slabs.rebalance(b, reduce, rank, world, cols)
        drv.step()
        info = b.last_info
        log.append([info[k] for k in ("n_total", "n_owned", "to_left", "to_right", "from_left", "from_right")] + [drv.resends])
    drv.drain()
    i, p, v, r = b.owned()
    counts.append(b.n_owned)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), ids=i, pos=p, vel=v, rho=r, log=np.array(log), counts=np.array(counts), cols=np.array(cols),
             transit=np.array([len(b.transit[0]) + len(b.transit[1]), b.forwarded]))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world,floor,lag", [(2, None, 0), (3, None, 0), (2, 1, 0), (2, None, 2), (3, None, 1)],
                         ids=["w2", "w3", "w2-resend", "w2-async2", "w3-async1"])
def test_slab_protocol_matches_single_domain(world, floor, lag):
    steps = 6
    port, slabs, P, G, cols = setup(world)
    pos, vel = scene()
    S = port.State(pos, vel)
    port.step_grid(P, G, S, steps)
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("spawn")
        pn = free_port()
        procs = [ctx.Process(target=worker, args=(r, world, pn, steps, d, 4096, floor, lag)) for r in range(world)]
        for p in procs: p.start()
        for p in procs: p.join(300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
    ids = np.concatenate([p["ids"] for p in parts])
    assert np.array_equal(np.sort(ids), np.arange(len(pos))), "every particle owned by exactly one slab"
    assert (S.pos[-3:, 0] > 0.5).all(), "the wall particles must have been clamped to the +x wall"
    o = np.argsort(ids)
    for name, want in (("pos", S.pos), ("vel", S.vel), ("rho", S.density)):
        got = np.concatenate([p[name] for p in parts])[o]
        assert np.array_equal(got, want), name
    logs = [p["log"] for p in parts]
    for r in range(world - 1):   # what r sends right is what r+1 receives from the left, every step
        assert np.array_equal(logs[r][:, 3], logs[r + 1][:, 4]) and np.array_equal(logs[r + 1][:, 2], logs[r][:, 5])
    assert sum(l[1:, 2:6].sum() for l in logs) > 0
    assert (sum(int(l[-1, 6]) for l in logs) > 0) == (floor is not None), "re-send path exercised only when sizes are forced too small"
    # migrations really happened: ownership after the run differs from the initial cell ownership
    moved = sum(int((l[:, 1][1:] != l[:, 1][:-1]).any()) for l in logs)
    assert moved > 0


def test_partition_columns():
    slabs = __import__("importlib").import_module("sph-erosion_b200.slabs")
    assert slabs.partition_columns(40, 4) == [(0, 10), (10, 20), (20, 30), (30, 40)]
    cols = slabs.partition_columns(100, 3, boundaries_x=[-0.5, 0.7], gmin_x=-2.0, cell=0.05)
    assert cols == [(0, 30), (30, 54), (54, 100)] or cols == [(0, 29), (29, 53), (53, 100)] or cols[0][0] == 0 and cols[-1][1] == 100
    with pytest.raises(ValueError):
        slabs.partition_columns(10, 4)


def test_far_migrant_is_forwarded_hop_by_hop():
    """A particle that crosses MORE than one slab in a step (slab 1 -> slab 3 of 4) is received by slab 2, which is not
    its owner, and handed on at the next exchange (slab.cu k_slab_append / k_slab_forward, restated in slab_double.py):
    owned + in transit is conserved every step, it ends up owned by slab 3, and every other particle stays bit-equal
    to the single-domain run."""
    world, steps = 4, 3
    port, slabs, P, G, cols = setup(world)
    pos, vel = scene()
    fast_p = np.array([[-0.2, 0.4, 0.0]], np.float32)       # isolated, in slab 1
    fast_v = np.array([[160.0, 0.0, 0.0]], np.float32)      # 0.64 box units per step: lands in slab 3
    n = len(pos) + 1
    S = port.State(np.concatenate([pos, fast_p]), np.concatenate([vel, fast_v]))
    port.step_grid(P, G, S, steps)
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("spawn")
        pn = free_port()
        procs = [ctx.Process(target=worker, args=(r, world, pn, steps, d, 4096, None, 0, (fast_p, fast_v))) for r in range(world)]
        for p in procs: p.start()
        for p in procs: p.join(300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
    ids = np.concatenate([p["ids"] for p in parts])
    in_transit = sum(int(p["transit"][0]) for p in parts)
    assert len(np.unique(ids)) == len(ids) and len(ids) + in_transit == n
    assert sum(int(p["transit"][1]) for p in parts) >= 1, "the scene must exercise forwarding"
    assert np.array_equal(np.sort(ids), np.arange(n)), "after 3 exchanges the fast particle is owned again"
    assert (n - 1) in parts[3]["ids"], "... by the slab its position belongs to"
    o = np.argsort(ids)
    got = np.concatenate([p["pos"] for p in parts])[o]
    assert np.array_equal(got[:n - 1], S.pos[:n - 1]), "everybody else is unaffected"


def test_rebalance_by_particle_count_keeps_results_bit_equal():
    """slabs.rebalance (SURVEY.md 8e "re-cut by particle-count quantiles"): an unbalanced scene (all particles in the left
    40 % of the box) on 3 equal-width slabs, re-cut before steps 1-4 (a cut may only move between its old neighbours, so a large imbalance converges over a
    few re-cuts).  The cuts move, the owned counts even out, every exchange migrates who changed owner, and the run stays BIT-EQUAL to the single-domain oracle."""
    world, steps = 3, 6
    port, slabs, P, G, cols = setup(world)
    pos, vel = scene()
    pos = squeeze_scene(pos, 0.4)
    S = port.State(pos, vel)
    port.step_grid(P, G, S, steps)
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("spawn")
        pn = free_port()
        procs = [ctx.Process(target=worker, args=(r, world, pn, steps, d, 4096, None, 0, None, (1, 2, 3, 4), 0.4)) for r in range(world)]
        for p in procs: p.start()
        for p in procs: p.join(300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
    ids = np.concatenate([p["ids"] for p in parts])
    assert np.array_equal(np.sort(ids), np.arange(len(pos)))
    o = np.argsort(ids)
    for name, want in (("pos", S.pos), ("vel", S.vel), ("rho", S.density)):
        assert np.array_equal(np.concatenate([p[name] for p in parts])[o], want), name
    before = [int(p["counts"][0]) for p in parts]; after = [int(p["counts"][-1]) for p in parts]
    assert max(before) > 0.65 * len(pos) and min(before) == 0, "the equal-width cut must be badly unbalanced: %r" % before
    assert max(after) < 0.5 * len(pos) and min(after) > 0.15 * len(pos), "the re-cuts even the load out: %r -> %r" % (before, after)
    new_cols = parts[0]["cols"].tolist()
    assert new_cols != [list(c) for c in cols] and all(np.array_equal(p["cols"], parts[0]["cols"]) for p in parts)
