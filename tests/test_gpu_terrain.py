"""GPU parity of the terrain half of the hot path, through the C ABI (Grid mirror -> libsphe_b200.so).

Bars:
  * Grid::collision (grid.h:462-805): hit/miss, contact point and normal BIT-EXACT against the fixtures
    produced by the unmodified reference (tests/golden/terrain.npz) and against the C oracle;
  * UpdateGrid / genIndices mesh (grid.h:118-176): BIT-EXACT;
  * terrain stage (contact response + this project's erosion model): BIT-EXACT against the oracle --
    integers (heights, sediment, hit flags) and floats (terrain.cu is built without FMA contraction);
  * full step with a terrain attached: fields within RTOL = 1e-5 of the oracle composition, same hit set;
  * conservation: sum(heights) + sum(carried sediment) is an EXACT integer invariant over many steps."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, product
from oracle import port

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "terrain.npz"))


def make_grid(gold, erosion=None):
    m = product()
    g = m.Grid(*[int(d) for d in gold["dims"]])
    img = np.zeros((512, 512), np.uint8); img[:64, :64] = gold["hf64"]
    g.LoadHeightfield(img)
    if erosion:
        e = g.erosion
        e.enabled = 1
        for k, v in erosion.items():
            setattr(e, k, v)
    return g


def oracle_terrain(gold):
    h = np.zeros((512, 512), np.float32); h[:64, :64] = gold["hf64"]
    return port.Terrain(h, gold["dims"])


def test_collision_bit_exact_vs_reference_fixture(gold):
    g = make_grid(gold)
    hit, cp, nrm = g.collision(gold["pc"], gold["pn"], gold["vn"])
    assert np.array_equal(hit, gold["hit"])
    assert np.array_equal(bits(cp), bits(gold["cp"]))
    assert np.array_equal(bits(nrm), bits(gold["nrm"]))


def test_collision_bit_exact_vs_oracle_random_field():
    rng = np.random.default_rng(21)
    m = product()
    h = (rng.integers(0, 30, (96, 80)) + 40 * (1 + np.cos(np.arange(80) / 4.0))[None, :]).astype(np.uint8)
    g = m.Grid(90, 255, 70); g.load_heightfield(h)
    T = port.Terrain(h.astype(np.float32), (90, 255, 70))
    n = 60000
    x = rng.uniform(2.5, 86.5, n); z = rng.uniform(2.5, 66.5, n)
    hl = h[np.floor(x).astype(int), np.floor(z).astype(int)].astype(np.float32)
    vd = rng.normal(0, 1, (n, 3)); vd[:, 1] = -np.abs(vd[:, 1]) * 2 - 0.2; vd /= np.linalg.norm(vd, axis=1)[:, None]
    pn = np.stack([x, hl - rng.uniform(0, 4, n) + rng.uniform(-3, 3, n), z], 1).astype(np.float32)
    pc = (pn - vd * np.minimum(rng.uniform(0.01, 2.5, n), 1.4)[:, None]).astype(np.float32)
    vn = (vd * rng.uniform(0.1, 5, n)[:, None]).astype(np.float32)
    h1, c1, n1 = g.collision(pc, pn, vn)
    h2, c2, n2 = T.collision(pc, pn, vn)
    assert 0.08 < h2.mean() < 0.5
    assert np.array_equal(h1, h2)
    assert np.array_equal(bits(c1), bits(c2)) and np.array_equal(bits(n1), bits(n2))


def test_mesh_bit_exact(gold):
    g = make_grid(gold)
    g.UpdateGrid(*[int(d) for d in gold["dims"]])
    assert g.GetSurfacePartsSize() == gold["surface"].size and g.GetIndicesSize() == gold["indices"].size
    assert np.array_equal(bits(g.GetSurfaceParts()), bits(gold["surface"]))
    assert np.array_equal(g.GetIndices(), gold["indices"])
    assert [g.GetHeightfieldAt(0, 0), g.GetHeightfieldAt(5, 7), g.GetHeightfieldAt(49, 49), g.GetHeightfieldAt(63, 1)] == gold["heights_probe"].tolist()
    assert g.GetDim() == (50, 255, 50)


@pytest.mark.parametrize("case", ["pickup", "deposit", "bedrock-share", "response-only", "scaled"])
def test_terrain_stage_bit_exact_vs_oracle(gold, case):
    h = np.flatnonzero(gold["hit"])
    miss = np.flatnonzero(gold["hit"] == 0)[:500]
    sel = np.concatenate([h, miss])
    pc, pn, vn = gold["pc"][sel].copy(), gold["pn"][sel].copy(), gold["vn"][sel].copy()
    n = len(sel)
    rng = np.random.default_rng(8)
    kw = dict(Kc=0.05, Ke=0.3, Kd=0.3, hmin=0.0, max_pickup=0.25)
    sed = np.zeros(n, np.int32)
    origin, scale = (0.0, 0.0, 0.0), 1.0
    if case == "deposit":
        sed = rng.integers(0, 6 * 4096, n).astype(np.int32)
    elif case == "bedrock-share":
        kw.update(hmin=float(gold["hf64"][1:48, 1:48].min()) + 20.0, Ke=4.0, Kc=4.0, max_pickup=40.0)
        pn = np.concatenate([pn, pn[:len(h)]]); pc = np.concatenate([pc, pc[:len(h)]]); vn = np.concatenate([vn, vn[:len(h)]])
        sed = np.zeros(len(pn), np.int32); n = len(pn)   # duplicated particles compete for the same vertices
    elif case == "scaled":
        origin, scale = (-0.16, -0.5, -0.16), 0.00625
        pc = (pc * np.float32(scale) + np.array(origin, np.float32)).astype(np.float32)
        pn = (pn * np.float32(scale) + np.array(origin, np.float32)).astype(np.float32)
        vn = (vn * np.float32(scale)).astype(np.float32)
        kw.update(Kc=8.0)
        sed = rng.integers(0, 300, n).astype(np.int32)
    enabled = case != "response-only"
    g = make_grid(gold, kw if enabled else None)
    g.set_transform(origin, scale)
    T = oracle_terrain(gold)
    E = port.erosion_params(enabled=enabled, origin=origin, scale=scale, **kw)
    p1, v1, s1 = pn.copy(), vn.copy(), sed.copy()
    p2, v2, s2 = pn.copy(), vn.copy(), sed.copy()
    tot0 = g.total_fx() + int(s1.astype(np.int64).sum())
    hit1 = g.stage(pc, p1, v1, s1, 0.01)
    hit2 = T.stage(E, pc, p2, v2, s2, 0.01)
    assert np.array_equal(hit1, hit2) and hit1.sum() >= 0.5 * len(h)
    assert np.array_equal(s1, s2), "carried sediment"
    assert np.array_equal(g.heights_fx(), T.hfx), "terrain heights"
    assert np.array_equal(bits(p1), bits(p2)) and np.array_equal(bits(v1), bits(v2))
    assert g.total_fx() + int(s1.astype(np.int64).sum()) == tot0, "exact conservation"
    if enabled:
        assert not np.array_equal(T.hfx, oracle_terrain(gold).hfx), "the case must change the terrain"
    if case == "bedrock-share":
        lim = int(round(kw["hmin"] * 4096))
        h0 = oracle_terrain(gold).hfx
        assert (g.heights_fx()[h0 >= lim] >= lim).all()


def _scene(n_side=14, seed=3):
    """A block of fluid resting on / falling into the lena patch: terrain cell = 0.0125 world units."""
    rng = np.random.default_rng(seed)
    i = np.arange(n_side)
    pos = np.stack(np.meshgrid(i, i[: n_side // 2], i, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.025
    pos += np.array([-0.17, -0.14, -0.17], np.float32) + rng.uniform(-0.003, 0.003, pos.shape).astype(np.float32)
    vel = np.zeros_like(pos); vel[:, 1] = -0.6; vel[:, 0] = 0.4
    return pos, vel


def _attach(gold, erosion):
    scale = 0.0125
    h = gold["hf64"][:40, :40].astype(np.float32)
    hh = (h - h.min()) / 8.0 + 2.0      # 2 .. ~28 cells high -> 0.025 .. 0.35 world units above the origin
    origin = (-0.25, -0.2, -0.25)
    m = product()
    g = m.Grid(40, 255, 40); g.set_heights(hh); g.set_transform(origin, scale)
    e = g.erosion
    e.enabled = int(erosion); e.Kc = 4.0; e.Ke = 0.5; e.Kd = 0.25; e.hmin = 1.0; e.max_pickup = 0.5
    T = port.Terrain(hh, (40, 255, 40))
    E = port.erosion_params(enabled=erosion, origin=origin, scale=scale, Kc=4.0, Ke=0.5, Kd=0.25, hmin=1.0, max_pickup=0.5)
    return g, T, E


@pytest.mark.parametrize("variant", [(6, 3), (0, 0), (20, 20)], ids=["default", "tpp", "staged"])
def test_step_with_terrain_lockstep_vs_oracle(gold, variant):
    m = product()
    pos, vel = _scene()
    g, T, E = _attach(gold, True)
    P = port.default_params(dt=0.004, len=0.3)
    s = m.FluidSystemSPH(); s.set_variant(*variant)
    q = s.params; q.dt = P.dt; q.len = P.len
    S = port.State(pos, vel)
    sed = np.zeros(S.n, np.int32)
    G = None
    hits = 0
    for step in range(12):
        s.upload_state(S.pos, S.vel)
        s.set_sediment_fx(sed)
        g.set_heights(T.h)              # lockstep: GPU starts every step from the oracle's state
        s.Run(g)
        if G is None:
            gi = s.grid_info(); G = port.Grid(); G.gmin[:] = list(gi.gmin); G.cell = gi.cell; G.dim[:] = list(gi.dim)
        hit = port.step_grid_terrain(P, G, S, T, E, sed)
        hits += int(hit.sum())
        for name, want in (("density", S.density), ("pos", S.pos), ("vel", S.vel)):
            got = s.download(name).astype(np.float64)
            scale = max(np.abs(want).max(), 1e-30)
            err = np.abs(got - want).max()
            assert err <= 20 * RTOL * scale, "step %d %s rel err %.2e" % (step, name, err / scale)
        # erosion amounts depend on float velocities that agree to ~1e-5: heights agree to a few fixed-point units
        dh = np.abs(g.heights_fx().astype(np.int64) - T.hfx.astype(np.int64))
        assert dh.max() <= 8, "step %d: height difference %d fixed-point units" % (step, dh.max())
    assert hits > 50, "the scene must exercise terrain contacts (%d)" % hits


def test_conservation_and_determinism_free_running(gold):
    m = product()
    pos, vel = _scene(n_side=20)
    runs = []
    for rep in range(2):
        g, T, E = _attach(gold, True)
        s = m.FluidSystemSPH()
        s.params.dt = 0.004; s.params.len = 0.3
        s.upload_state(pos, vel)
        tot0 = g.total_fx() + s.sediment_total_fx()
        h0 = g.heights_fx()
        for step in range(60):
            s.Run(g)
            if step % 10 == 9:
                assert g.total_fx() + s.sediment_total_fx() == tot0, "step %d" % step
        h1 = g.heights_fx()
        assert (h1 != h0).sum() > 20 and s.sediment_total_fx() > 0
        assert h1.min() >= min(int(h0.min()), 4096), "never below bedrock"
        p = s.download("pos")
        # the reference box clamps only the axis of largest |coordinate| per step (fluid_system.h:362-371),
        # so corners overshoot slightly (+-0.207 for len 0.2)
        assert np.isfinite(p).all() and np.abs(p).max() <= 0.3 * 1.05
        runs.append((h1, s.download("sediment"), p))
    # integer atomics: the terrain is identical run to run
    assert np.array_equal(runs[0][0], runs[1][0])
    assert np.array_equal(bits(runs[0][2]), bits(runs[1][2]))


def test_no_terrain_and_flat_far_terrain_agree():
    """A terrain far below the fluid must not change the step (culling by the max height is exact)."""
    m = product()
    pos, vel = _scene()
    a = m.FluidSystemSPH(); b = m.FluidSystemSPH()
    for s in (a, b):
        s.params.dt = 0.004; s.params.len = 0.3
        s.upload_state(pos, vel)
    g = m.Grid(40, 255, 40); g.set_heights(np.zeros((40, 40), np.float32)); g.set_transform((-0.25, -5.0, -0.25), 0.0125)
    for _ in range(5):
        a.Run(); b.Run(g)
    assert np.array_equal(bits(a.download("pos")), bits(b.download("pos")))
    assert np.array_equal(bits(a.download("vel")), bits(b.download("vel")))


def test_checkpoint_resume_continues_bit_for_bit(tmp_path):
    """sphe_save_state / sphe_load_state: particles (+ carried sediment, fixed point) and the eroded terrain go to one
    file; a fresh handle + a fresh terrain loaded from it continue EXACTLY like the uninterrupted run -- positions,
    velocities, densities, sediment, heights, contact for contact.  Also: a file with a terrain refuses to load
    without a terrain handle, and a non-state file is rejected."""
    import importlib
    m = product()
    T = importlib.import_module("test_gpu_slabs")
    box = (1.2, 0.3, 0.3)
    params = dict(len=0.3, dt=0.004, g=(0.0, -9.82, 0.0))
    g1, pos, vel = T._terrain_scene(m)
    one = T._single(m, box, params, (6, 3), pos, vel)
    for _ in range(6):
        one.Run(g1)
    ck = str(tmp_path / "state.sphe")
    one.save_state(ck, g1)
    for _ in range(6):
        one.Run(g1)

    two = m.FluidSystemSPH()
    g2 = m.Grid(4, 255, 4)                      # wrong size on purpose: the file brings its own
    with pytest.raises(m.capi.SpheError, match="terrain"):
        two.load_state(ck)
    two.load_state(ck, g2)
    assert g2.shape() == g1.shape() and two.count() == one.count()
    for _ in range(6):
        two.Run(g2)
    for f in ("pos", "vel", "density", "sediment"):
        assert np.array_equal(one.download(f), two.download(f)), f
    assert np.array_equal(g1.heights_fx(), g2.heights_fx())
    assert one.sediment_total_fx() == two.sediment_total_fx() > 0

    bad = tmp_path / "bad.sphe"; bad.write_bytes(b"not a state file at all" * 10)
    with pytest.raises(m.capi.SpheError, match="not a sphe state file"):
        two.load_state(str(bad))


def test_load_state_rejects_corrupt_files_without_touching_the_simulation(tmp_path):
    """A truncated file, a header that promises more particles than the file holds, an oversized terrain: each is an
    error and leaves the simulation and the terrain exactly as they were (sphe_load_state reads and checks the whole
    file before it mutates anything); saving never leaves a partial file behind."""
    m = product()
    s = m.FluidSystemSPH(); s.params.dt = 0.004; s.params.len = 0.3
    pos = np.random.default_rng(2).uniform(-0.25, 0.25, (5000, 3)).astype(np.float32)
    s.upload_state(pos, np.zeros_like(pos))
    g = m.Grid(40, 255, 40); g.set_heights(np.full((40, 40), 3.0, np.float32)); g.set_transform((-0.3, -0.3, -0.3), 0.015)
    s.Run(g)
    good = tmp_path / "good.sphe"
    s.save_state(str(good), g)
    assert not (tmp_path / "good.sphe.tmp").exists()
    raw = good.read_bytes()
    p0, h0, n0 = s.download("pos"), g.heights_fx(), s.count()
    cases = {"truncated_particles": raw[:len(raw) // 3], "truncated_terrain": raw[:-1000]}
    huge = bytearray(raw); huge[8:12] = (2 ** 30).to_bytes(4, "little")      # n = 2^30 particles in a 200 KB file
    cases["huge_n"] = bytes(huge)
    for name, data in cases.items():
        f = tmp_path / (name + ".sphe"); f.write_bytes(data)
        with pytest.raises(m.capi.SpheError):
            s.load_state(str(f), g)
        assert s.count() == n0 and np.array_equal(s.download("pos"), p0) and np.array_equal(g.heights_fx(), h0), name
    s.load_state(str(good), g)
    assert np.array_equal(s.download("pos"), p0) and np.array_equal(g.heights_fx(), h0)
    with pytest.raises(m.capi.SpheError):
        s.save_state(str(tmp_path / "no_such_dir" / "x.sphe"), g)
