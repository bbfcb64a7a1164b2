"""Terrain oracle (oracle/terrain_oracle.c) against (a) the committed fixtures produced by the
unmodified reference on lena_gray.png (tests/golden/terrain.npz, made by make_golden.py terrain) and
(b) the compiled reference itself when oracle/_ref is present.  Bar: BIT-EXACT hit/miss decisions,
contact points, normals, mesh vertices/normals and indices."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import port, ref


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "terrain.npz"))


def terrain_from(gold):
    return port.Terrain(gold["hf64"].astype(np.float32), gold["dims"])


def test_collision_matches_reference_fixture(gold):
    T = terrain_from(gold)
    hit, cp, nrm = T.collision(gold["pc"], gold["pn"], gold["vn"])
    assert np.array_equal(hit, gold["hit"])
    assert 0.1 < hit.mean() < 0.3
    assert np.array_equal(bits(cp), bits(gold["cp"]))
    assert np.array_equal(bits(nrm), bits(gold["nrm"]))
    # the contact point lies on the heightfield surface facet it was mapped to (1e-5)
    h = np.flatnonzero(hit)
    n = nrm[h]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5) and (n[:, 1] > 0).all()


def test_mesh_matches_reference_fixture(gold):
    T = terrain_from(gold)
    assert np.array_equal(bits(T.surface()), bits(gold["surface"]))
    assert np.array_equal(T.indices(), gold["indices"])
    hf = gold["hf64"]
    assert [int(hf[0, 0]), int(hf[5, 7]), int(hf[49, 49]), int(hf[63, 1])] == gold["heights_probe"].tolist()
    assert gold["voxel_probe"].tolist() == [[2, 0], [2, 0]]  # VOXEL_MAT at the surface voxel, AIR below (voxel.h:1-6)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_collision_random_vs_compiled_reference():
    rng = np.random.default_rng(99)
    # a synthetic rough field: the reference Grid always owns a 512 x 512 heightfield (grid.h:78-81)
    img = (rng.integers(0, 40, (512, 512)) + 60 * (1 + np.sin(np.arange(512) / 3.0))[:, None]).astype(np.uint8)
    g = ref.RefGrid(80, 255, 80); g.load_heightfield(img); g.update(80, 255, 80)
    T = port.Terrain(img.astype(np.float32), (80, 255, 80))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    for pen, step in [(0.05, 0.1), (1.0, 1.0), (6.0, 3.0)]:
        pc, pn, vn = mg.terrain_trials(rng, img, 20000, pen, step, lo=2.5, hi=76.5)
        h1, c1, n1 = g.collision(pc, pn, vn)
        h2, c2, n2 = T.collision(pc, pn, vn)
        assert np.array_equal(h1, h2)
        assert np.array_equal(bits(c1), bits(c2)) and np.array_equal(bits(n1), bits(n2))
    assert np.array_equal(bits(g.surface()), bits(T.surface()))
    assert np.array_equal(g.indices(), T.indices())


def test_erosion_stage_invariants():
    """The erosion model is this project's own (no reference code): exact conservation, monotone
    pure-deposit / pure-pickup, nothing happens at dt = 0 or when erosion is disabled."""
    gold = np.load(os.path.join(GOLDEN, "terrain.npz"))
    rng = np.random.default_rng(4)
    hit0 = np.flatnonzero(gold["hit"])[:1000]
    pc, pn, vn = gold["pc"][hit0].copy(), gold["pn"][hit0].copy(), gold["vn"][hit0].copy()

    def run(E, sed0, dt=0.01):
        T = terrain_from(gold)
        sed = sed0.copy(); p = pn.copy(); v = vn.copy()
        hit = T.stage(E, pc, p, v, sed, dt)
        return T, sed, p, v, hit

    zero = np.zeros(len(hit0), np.int32)
    T0 = terrain_from(gold)
    total0 = int(T0.hfx.astype(np.int64).sum())
    # pick-up only (particles carry nothing)
    T, sed, p, v, hit = run(port.erosion_params(), zero)
    assert hit.all()
    assert (sed >= 0).all() and sed.sum() > 0
    assert int(T.hfx.astype(np.int64).sum()) + int(sed.astype(np.int64).sum()) == total0, "exact conservation"
    assert (T.hfx <= T0.hfx).all(), "pure pick-up never raises the terrain"
    assert np.array_equal(T.h, T.hfx.astype(np.float32) / np.float32(4096))
    # deposit only (particles saturated)
    full = np.full(len(hit0), 8 * 4096, np.int32)
    T, sed, _, _, _ = run(port.erosion_params(), full)
    assert int(T.hfx.astype(np.int64).sum()) + int(sed.astype(np.int64).sum()) == total0 + int(full.astype(np.int64).sum())
    assert (T.hfx >= T0.hfx).all() and (sed <= full).all() and (sed >= 0).all()
    # bedrock: nothing can be taken below hmin
    E = port.erosion_params(hmin=1000.0)
    T, sed, _, _, _ = run(E, zero)
    assert sed.sum() == 0 and np.array_equal(T.hfx, T0.hfx)
    # a thin layer above bedrock is shared out, never overdrawn
    hmin = float(T0.h.min()) + 0.01
    T, sed, _, _, _ = run(port.erosion_params(hmin=hmin, Ke=5.0, Kc=5.0, max_pickup=50.0), zero)
    assert (T.hfx >= min(int(T0.hfx.min()), int(round(hmin * 4096)))).all()
    assert (T.hfx[T0.hfx >= int(round(hmin * 4096))] >= int(round(hmin * 4096))).all()
    assert int(T.hfx.astype(np.int64).sum()) + int(sed.astype(np.int64).sum()) == total0
    # disabled / paused: response only, or nothing at all
    T, sed, p, v, hit = run(port.erosion_params(enabled=False), zero)
    assert hit.all() and sed.sum() == 0 and np.array_equal(T.hfx, T0.hfx)
    assert np.array_equal(bits(p), bits(gold["cp"][hit0])), "position snaps to the reference's contact point"
    T, sed, p, v, hit = run(port.erosion_params(), zero, dt=0.0)
    assert not hit.any() and np.array_equal(bits(p), bits(pn)) and np.array_equal(bits(v), bits(vn))
