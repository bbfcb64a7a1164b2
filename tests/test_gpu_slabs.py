"""Multi-GPU slab kernels (slab.cu) on ONE GPU: K slab handles driven in one process
(slabs.LocalSlabGroup: the same pack/unpack/step calls the NCCL driver makes, the P2P replaced
by reading the neighbour's send buffer) must reproduce the single-handle run.

Bar: BIT-EQUAL positions, velocities and densities for every particle after every step -- a slab sees
each owned particle's neighbours in the same canonical (cell, id) order as the single-GPU run, ghosts
included (SURVEY.md 8e "1-GPU vs k-GPU bit-equal")."""
import importlib

import numpy as np
import pytest

from conftest import product

pytestmark = pytest.mark.gpu


def scene(n=20000, seed=11):
    rng = np.random.default_rng(seed)
    pos = np.empty((n, 3), np.float32)
    pos[:, 0] = rng.uniform(-1.1, 1.1, n); pos[:, 1] = rng.uniform(-0.29, 0.0, n); pos[:, 2] = rng.uniform(-0.29, 0.29, n)
    vel = rng.normal(0, 1.5, (n, 3)).astype(np.float32)
    # isolated particles EXACTLY on the -x wall with no x velocity: the reference's box response clamps them to the
    # +x wall (collisionS, fluid_system.h:375-382) -- from the first slab straight to the last one (sphe_slab_ring)
    wall = np.array([[-1.2, 0.1 + 0.06 * k, -0.2 + 0.1 * k] for k in range(4)], np.float32)
    pos = np.concatenate([pos, wall]); vel = np.concatenate([vel, np.zeros_like(wall)])
    return pos, vel


@pytest.mark.parametrize("async_", [False, True], ids=["sync", "async"])
@pytest.mark.parametrize("variant", [(6, 3), (0, 0), (20, 20)], ids=["default", "tpp", "staged"])
@pytest.mark.parametrize("K", [2, 3])
def test_k_slabs_equal_single_gpu(K, variant, async_):
    import torch
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    pos, vel = scene()
    n = pos.shape[0]
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))

    one = m.FluidSystemSPH()
    for k, v in params.items():
        if k == "g": one.params.g[0], one.params.g[1], one.params.g[2] = v
        else: setattr(one.params, k, v)
    one.set_box(box); one.set_variant(*variant)
    one.upload_state(pos, vel)

    sims, backs = [], []
    for r in range(K):
        sim, b, cols = slabs.make_gpu_slab(m, torch.cuda.current_device(), r, K, box, params, None, 1 << 15, variant)
        sims.append(sim); backs.append(b)
    # initial distribution by x quantile -- not by cell column, so the first exchange has to migrate
    order = np.argsort(pos[:, 0], kind="stable")
    for r, part in enumerate(np.array_split(order, K)):
        sims[r].slab_upload(pos[part], vel[part], part.astype(np.int32))
    group = slabs.LocalSlabGroup(backs, async_=async_)

    owners_prev = None
    migrated = 0
    for step in range(8):
        one.Run()
        group.step()
        if async_ and step % 3 != 2:
            continue          # let the host run ahead: no sync, no download for a few steps
        if async_:
            group.drain()
        got = [s.slab_download() for s in sims]
        ids = np.concatenate([g[0] for g in got])
        assert np.array_equal(np.sort(ids), np.arange(n)), "step %d: every particle owned exactly once" % step
        o = np.argsort(ids)
        for j, name in ((1, "pos"), (2, "vel"), (3, "density")):
            a = np.concatenate([g[j] for g in got])[o]
            b = one.download(name)
            assert np.array_equal(a, b), "step %d %s: max |diff| %.3e" % (step, name, np.abs(a - b).max())
        owners = np.concatenate([np.full(len(g[0]), r) for r, g in enumerate(got)])[o]
        if owners_prev is not None:
            migrated += int((owners != owners_prev).sum())
        owners_prev = owners
    assert migrated > 0, "the scene must exercise migration"
    info = [s.slab_info() for s in sims]
    assert sum(i["n_owned"] for i in info) == n
    assert all(i["n_total"] > i["n_owned"] for i in info), "every slab holds ghosts"


def test_anisotropic_box_matches_oracle_free_particles():
    """sphe_set_box: walls at per-axis half-extents (no neighbours: h tiny so particles are free)."""
    m = product()
    s = m.FluidSystemSPH()
    s.params.dt = 0.01; s.params.len = 0.3
    s.set_box((1.0, 0.3, 0.5))
    pos = np.array([[0.995, 0.0, 0.0], [0.0, -0.299, 0.0], [0.0, 0.0, 0.499], [0.5, 0.1, 0.2]], np.float32)
    vel = np.array([[2.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, 3.0], [0.1, 0.1, 0.1]], np.float32)
    s.upload_state(pos, vel)
    s.Run()
    p = s.download("pos")
    assert p[0, 0] == np.float32(1.0) and p[1, 1] == np.float32(-0.3) and p[2, 2] == np.float32(0.5)
    assert abs(p[3, 0] - 0.501) < 1e-3
    v = s.download("vel")
    assert v[0, 0] < 0 and v[2, 2] < 0


# --------------------------------------------------------------------------- peer-memory exchange
def _single(m, box, params, variant, pos, vel):
    one = m.FluidSystemSPH()
    for k, v in params.items():
        if k == "g": one.params.g[0], one.params.g[1], one.params.g[2] = v
        else: setattr(one.params, k, v)
    one.set_box(box); one.set_variant(*variant)
    one.upload_state(pos, vel)
    return one


def _peer_group(m, slabs, K, box, params, variant, pos, vel, grid=None, cap=1 << 15):
    import torch
    sims = []
    for r in range(K):
        sim, b, cols = slabs.make_gpu_slab(m, torch.cuda.current_device(), r, K, box, params, None, cap, variant)
        sims.append(sim)
    order = np.argsort(pos[:, 0], kind="stable")
    for r, part in enumerate(np.array_split(order, K)):
        sims[r].slab_upload(pos[part], vel[part], part.astype(np.int32))
    return sims, slabs.LocalPeerGroup(sims, cap, pos.shape[0] + 4 * cap, grid=grid)


def _gather(sims, n):
    got = [s.slab_download() for s in sims]
    ids = np.concatenate([g[0] for g in got])
    assert np.array_equal(np.sort(ids), np.arange(n)), "every particle owned exactly once"
    o = np.argsort(ids)
    return [np.concatenate([g[j] for g in got])[o] for j in range(1, 5)]   # pos, vel, density, sediment (float bits)


@pytest.mark.parametrize("K", [2, 3, 5])
def test_peer_mailbox_slabs_equal_single_gpu(K):
    """k_slab_classify<true> storing into the neighbours' mailboxes + k_slab_append<true> waiting on the flags
    (the multi-GPU default, here with all slabs in one process): bit-equal to the single-handle run, with
    the host never syncing between steps."""
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    pos, vel = scene()
    n = pos.shape[0]
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    one = _single(m, box, params, (6, 3), pos, vel)
    sims, group = _peer_group(m, slabs, K, box, params, (6, 3), pos, vel)
    for step in range(9):
        one.Run()
        group.step()
        if step % 3 != 2:
            continue
        info = group.drain()
        assert all(i["from_left"] + i["from_right"] > 0 for i in info)
        p, v, rho, _ = _gather(sims, n)
        for a, name in ((p, "pos"), (v, "vel"), (rho, "density")):
            b = one.download(name)
            assert np.array_equal(a, b), "step %d %s: max |diff| %.3e" % (step, name, np.abs(a - b).max())
    assert (p[-4:, 0] > 1.0).all(), "the particles that started on the -x wall were clamped to the +x wall (ring closure)"


def test_peer_mailbox_overflow_is_reported():
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    pos, vel = scene()
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    sims, group = _peer_group(m, slabs, 2, box, params, (6, 3), pos, vel, cap=64)   # far too small for the halo
    group.step()
    with pytest.raises(m.capi.SpheError, match="overflow"):
        group.drain()


def test_peer_recv_times_out_instead_of_hanging():
    """A consumer whose neighbour never sends gives up after the timeout and reports it."""
    import torch
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    pos, vel = scene(2000)
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    sims, group = _peer_group(m, slabs, 2, box, params, (6, 3), pos, vel)
    sims[0].slab_peer_timeout(20_000_000)   # ~10 ms
    sims[0].slab_send()                     # slab 1 never sends
    t = sims[0].slab_recv()
    with pytest.raises(m.capi.SpheError, match="did not arrive"):
        sims[0].slab_result(t, True)


def _local_top(hts, pos, cell):
    rows, cols = hts.shape
    ix = np.clip(((pos[:, 0] + 1.2) / cell).astype(int), 0, rows - 2); iz = np.clip(((pos[:, 2] + 0.3) / cell).astype(int), 0, cols - 2)
    return np.maximum.reduce([hts[ix, iz], hts[ix + 1, iz], hts[ix, iz + 1], hts[ix + 1, iz + 1]])


def _terrain_scene(m, n=30000, seed=5):
    """Particles raining on a rough 256 x 64 terrain strip that fills the floor of a (1.2, 0.3, 0.3) box."""
    rng = np.random.default_rng(seed)
    rows, cols = 256, 64
    hts = (rng.uniform(0, 6, (rows, cols)) + 4 * np.sin(np.arange(rows) / 9.0)[:, None] + 6).astype(np.float32)
    g = m.Grid(rows, 255, cols)
    g.set_heights(hts)
    cell = 2.4 / rows
    g.set_transform((-1.2, -0.3, -0.3), cell)
    e = g.erosion
    e.enabled = 1; e.Kc = 2.0; e.Ke = 0.4; e.Kd = 0.3; e.hmin = 2.0; e.max_pickup = 0.5
    pos = np.empty((n, 3), np.float32)
    pos[:, 0] = rng.uniform(-1.15, 1.15, n); pos[:, 2] = rng.uniform(-0.28, 0.28, n)
    ix = np.clip(((pos[:, 0] + 1.2) / cell).astype(int), 0, rows - 2); iz = np.clip(((pos[:, 2] + 0.3) / cell).astype(int), 0, cols - 2)
    local_top = np.maximum.reduce([hts[ix, iz], hts[ix + 1, iz], hts[ix, iz + 1], hts[ix + 1, iz + 1]])
    pos[:, 1] = -0.3 + local_top * cell + rng.uniform(0.001, 0.06, n)      # just above the local surface
    vel = rng.normal(0, 0.6, (n, 3)).astype(np.float32); vel[:, 1] -= 1.0
    return g, pos, vel


@pytest.mark.parametrize("K", [2, 3])
def test_slabs_sharing_one_eroding_terrain_equal_single_gpu(K):
    """Erosion across slab boundaries: requests of owned particles only, per-vertex sums over all slabs
    between the phases (sphe_step_phase).  Heights, carried sediment and particle state must be BIT-EQUAL
    to the single-handle run, and sum(heights) + sum(sediment) exactly conserved."""
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    g1, pos, vel = _terrain_scene(m)
    gk, _, _ = _terrain_scene(m)
    n = pos.shape[0]
    one = _single(m, box, params, (6, 3), pos, vel)
    sims, group = _peer_group(m, slabs, K, box, params, (6, 3), pos, vel, grid=gk)
    total0 = g1.total_fx()
    for step in range(12):
        one.Run(g1)
        group.step()
    group.drain()
    assert g1.contacts() > 1000, "the scene must exercise terrain contacts"
    assert g1.contacts() == gk.contacts()
    assert np.array_equal(g1.heights_fx(), gk.heights_fx())
    p, v, rho, sed = _gather(sims, n)
    assert np.array_equal(p, one.download("pos")) and np.array_equal(v, one.download("vel"))
    assert np.array_equal(rho, one.download("density"))
    sed_k = sum(s.sediment_total_fx() for s in sims)
    assert sed_k == one.sediment_total_fx() and sed_k > 0
    assert gk.total_fx() + sed_k == total0, "sum(heights) + sum(carried sediment) is conserved exactly"


@pytest.mark.parametrize("K", [2, 3])
def test_slab_local_terrain_windows_equal_single_gpu(K):
    """The multi-GPU default for the terrain (slabs.TerrainWindowShare): every slab has its OWN terrain replica and
    keeps only the rows under the slab + a margin current (sphe_terrain_set_window); the erosion accumulators are
    summed with the x-neighbour over the rows around the common boundary only.  Every row a slab owns, the
    carried sediment and the particle state must be BIT-EQUAL to the single-handle run; no contact may reach
    outside a window; sum(owned rows) + sum(sediment) is conserved exactly."""
    import torch
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    g1, pos, vel = _terrain_scene(m)
    n = pos.shape[0]
    one = _single(m, box, params, (6, 3), pos, vel)
    total0 = g1.total_fx()
    cap = 1 << 15
    sims, replicas = [], []
    for r in range(K):
        sim, b, cols = slabs.make_gpu_slab(m, torch.cuda.current_device(), r, K, box, params, None, cap, (6, 3))
        sims.append(sim); replicas.append(_terrain_scene(m)[0])
    order = np.argsort(pos[:, 0], kind="stable")
    for r, part in enumerate(np.array_split(order, K)):
        sims[r].slab_upload(pos[part], vel[part], part.astype(np.int32))
    gi = sims[0].grid_info()
    cell_t = 2.4 / 256
    cuts = slabs.terrain_row_cuts(gi, cols, -1.2, cell_t)
    W = slabs.terrain_margin_rows(gi.cell, cell_t)
    dev = torch.device("cuda", torch.cuda.current_device())
    shares = [slabs.TerrainWindowShare(replicas[r], dev, r, K, cuts, W, swap=False) for r in range(K)]
    assert all(s.window[1] - s.window[0] < 256 for s in shares), "the windows must be real restrictions"
    group = slabs.LocalPeerGroup(sims, cap, n + 4 * cap, shares=shares)
    for step in range(12):
        one.Run(g1)
        group.step()
    group.drain()
    assert g1.contacts() > 1000
    assert sum(g.contacts() for g in replicas) == g1.contacts()
    assert all(g.window_violations() == 0 for g in replicas)
    want = g1.heights_fx()
    assert not np.array_equal(want, _terrain_scene(m)[0].heights_fx()), "the terrain must have eroded"
    for r, (g, sh) in enumerate(zip(replicas, shares)):
        got = g.heights_fx()
        assert np.array_equal(got[sh.own[0]:sh.own[1]], want[sh.own[0]:sh.own[1]]), "slab %d: owned terrain rows" % r
        assert np.array_equal(got[sh.window[0]:sh.window[1]], want[sh.window[0]:sh.window[1]]), "slab %d: window rows" % r
    p, v, rho, sed = _gather(sims, n)
    assert np.array_equal(p, one.download("pos")) and np.array_equal(v, one.download("vel"))
    assert np.array_equal(rho, one.download("density"))
    sed_k = sum(s.sediment_total_fx() for s in sims)
    assert sed_k == one.sediment_total_fx() and sed_k > 0
    assert sum(sh.own_total_fx() for sh in shares) + sed_k == total0, "sum(owned rows) + sum(carried sediment) is conserved exactly"


def test_particles_crossing_more_than_one_slab_are_forwarded_not_lost():
    """A particle that lands beyond its neighbour slab (the reference's contact response can eject particles that
    fast) is handed on by the slab that received it, one exchange per extra slab (sphe_slab_transit): nothing is
    lost, nothing is duplicated, and every other particle stays bit-equal to the single-handle run."""
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    pos, vel = scene()
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    # isolated, far above the fluid: +x at 150 units/s = 13 columns per step (a slab of K = 5 is 11 columns wide), and -x
    fast = np.array([[-1.0, 0.15, 0.0], [-0.9, 0.2, 0.1], [1.0, 0.25, -0.1]], np.float32)
    fvel = np.array([[150.0, 0, 0], [230.0, 0, 0], [-260.0, 0, 0]], np.float32)
    pos = np.concatenate([pos, fast]); vel = np.concatenate([vel, fvel])
    n = pos.shape[0]
    one = _single(m, box, params, (6, 3), pos, vel)
    sims, group = _peer_group(m, slabs, 5, box, params, (6, 3), pos, vel)
    seen_transit = 0
    for step in range(8):
        one.Run()
        group.step()
        group.drain()
        got = [s.slab_download() for s in sims]
        ids = np.concatenate([g[0] for g in got])
        in_transit = sum(t["to_left"] + t["to_right"] for t in (s.slab_transit() for s in sims))
        seen_transit += in_transit
        assert len(np.unique(ids)) == len(ids), "step %d: a particle is owned twice" % step
        assert len(ids) + in_transit == n, "step %d: %d owned + %d in transit != %d" % (step, len(ids), in_transit, n)
    assert seen_transit > 0 and sum(s.slab_transit()["forwarded"] for s in sims) > 0, "the scene must exercise forwarding"
    assert np.array_equal(np.sort(ids), np.arange(n)), "after the fast particles hit the walls everybody is owned again"
    o = np.argsort(ids)
    p = np.concatenate([g[1] for g in got])[o]
    assert np.array_equal(p[:n - 3], one.download("pos")[:n - 3]), "the rest of the scene is unaffected"


def test_rebalance_on_the_gpu_keeps_results_bit_equal():
    """slabs.rebalance_local on real slab handles: an unbalanced cloud on 3 equal-width slabs, re-cut by particle count
    before steps 2-5 (slab_download histogram, slab_configure, slab_ring -- mid-run); the exchange migrates whoever
    changed owner and the run stays BIT-EQUAL to the single-handle run while the owned counts even out."""
    import torch
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    pos, vel = scene()
    pos = pos.copy(); pos[:-4, 0] = (-1.1 + (pos[:-4, 0] + 1.1) * np.float32(0.4)).astype(np.float32)   # everybody into the left 40 %
    n = pos.shape[0]
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    one = _single(m, box, params, (6, 3), pos, vel)
    K, cap = 3, 1 << 15
    sims, backs = [], []
    for r in range(K):
        sim, b, cols = slabs.make_gpu_slab(m, torch.cuda.current_device(), r, K, box, params, None, cap, (6, 3))
        sims.append(sim); backs.append(b)
    # initial distribution by OWNER (cell column): with x-quantile thirds the right third of this squeezed cloud would
    # start two slabs away from its owner and be forwarded, i.e. arrive a step late
    gi = sims[0].grid_info()
    cx = np.clip(np.floor((pos[:, 0] - np.float32(gi.gmin[0])) / np.float32(gi.cell)), 0, cols[-1][1] - 1).astype(np.int64)
    owner = np.searchsorted([c[1] for c in cols], cx, side="right")
    for r in range(K):
        part = np.nonzero(owner == r)[0]
        sims[r].slab_upload(pos[part], vel[part], part.astype(np.int32))
    group = slabs.LocalPeerGroup(sims, cap, n + 4 * cap)
    counts = []
    for step in range(8):
        if step in (2, 3, 4, 5):
            group.drain()
            counts.append([s.slab_info()["n_owned"] for s in sims])
            cols = slabs.rebalance_local(backs, cols)
        one.Run()
        group.step()
    group.drain()
    counts.append([s.slab_info()["n_owned"] for s in sims])
    p, v, rho, _ = _gather(sims, n)
    for a, name in ((p, "pos"), (v, "vel"), (rho, "density")):
        assert np.array_equal(a, one.download(name)), name
    assert max(counts[0]) > 0.6 * n and max(counts[-1]) < 0.5 * n and min(counts[-1]) > 0.15 * n, counts


def test_rebalance_with_slab_local_terrain_windows_keeps_results_bit_equal():
    """Re-cuts with a shared eroding terrain (TerrainWindowShare.recut): an unbalanced rain on 3 equal-width slabs, each with
    its own terrain replica and row window; before steps 3, 5 and 7 the slabs are re-cut by particle count, every row is
    brought up to date from its owner, the windows move with the cuts and the cull maps are rebuilt.  Particle state,
    carried sediment and every owned / window terrain row stay BIT-EQUAL to the single-handle run; conservation is exact."""
    import torch
    m = product()
    slabs = importlib.import_module("sph-erosion_b200.slabs")
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    g1, pos, vel = _terrain_scene(m)
    # squeeze the rain into the left 45 % of the strip (each particle keeps its height above the LOCAL surface)
    hts = g1.heights()
    cell_t = 2.4 / 256
    above = pos[:, 1] - (-0.3 + _local_top(hts, pos, cell_t) * cell_t)
    pos = pos.copy(); pos[:, 0] = (-1.15 + (pos[:, 0] + 1.15) * np.float32(0.45)).astype(np.float32)
    pos[:, 1] = (-0.3 + _local_top(hts, pos, cell_t) * cell_t + above).astype(np.float32)
    n = pos.shape[0]
    one = _single(m, box, params, (6, 3), pos, vel)
    total0 = g1.total_fx()
    K, cap = 3, 1 << 15
    sims, backs, replicas = [], [], []
    for r in range(K):
        sim, b, cols = slabs.make_gpu_slab(m, torch.cuda.current_device(), r, K, box, params, None, cap, (6, 3))
        sims.append(sim); backs.append(b); replicas.append(_terrain_scene(m)[0])
    gi = sims[0].grid_info()
    cx = np.clip(np.floor((pos[:, 0] - np.float32(gi.gmin[0])) / np.float32(gi.cell)), 0, cols[-1][1] - 1).astype(np.int64)
    owner = np.searchsorted([c[1] for c in cols], cx, side="right")
    for r in range(K):
        part = np.nonzero(owner == r)[0]
        sims[r].slab_upload(pos[part], vel[part], part.astype(np.int32))
    W = slabs.terrain_margin_rows(gi.cell, cell_t)
    dev = torch.device("cuda", torch.cuda.current_device())
    shares = [slabs.TerrainWindowShare(replicas[r], dev, r, K, slabs.terrain_row_cuts(gi, cols, -1.2, cell_t), W, swap=False)
              .bind_columns(gi, -1.2, cell_t) for r in range(K)]
    group = slabs.LocalPeerGroup(sims, cap, n + 4 * cap, shares=shares)
    counts, windows = [], [tuple(sh.window for sh in shares)]
    for step in range(10):
        if step in (3, 5, 7):
            group.drain()
            counts.append([s.slab_info()["n_owned"] for s in sims])
            cols = slabs.rebalance_local(backs, cols, min_width=shares[0].min_columns())
            slabs.TerrainWindowShare.recut_local(shares, shares[0].rows_of(cols))
            windows.append(tuple(sh.window for sh in shares))
        one.Run(g1)
        group.step()
    group.drain()
    counts.append([s.slab_info()["n_owned"] for s in sims])
    assert len(set(windows)) > 1, "the windows must have moved"
    assert g1.contacts() > 1000 and sum(g.contacts() for g in replicas) == g1.contacts()
    assert all(g.window_violations() == 0 for g in replicas)
    want = g1.heights_fx()
    assert not np.array_equal(want, _terrain_scene(m)[0].heights_fx()), "the terrain must have eroded"
    for r, (g, sh) in enumerate(zip(replicas, shares)):
        got = g.heights_fx()
        assert np.array_equal(got[sh.window[0]:sh.window[1]], want[sh.window[0]:sh.window[1]]), "slab %d: window rows" % r
    p, v, rho, sed = _gather(sims, n)
    assert np.array_equal(p, one.download("pos")) and np.array_equal(v, one.download("vel"))
    assert np.array_equal(rho, one.download("density"))
    sed_k = sum(s.sediment_total_fx() for s in sims)
    assert sed_k == one.sediment_total_fx() and sed_k > 0
    assert sum(sh.own_total_fx() for sh in shares) + sed_k == total0, "sum(owned rows) + sum(carried sediment) is conserved exactly"
    assert max(counts[-1]) < max(counts[0]), counts
