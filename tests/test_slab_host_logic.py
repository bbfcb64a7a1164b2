"""Host-side logic of the multi-GPU driver that needs no GPU: ring topology, terrain row cuts / margins, and the
zone layout of the slab-local terrain (slabs.TerrainWindowShare) checked with a stub terrain: the rows two
neighbours exchange must be the SAME rows, windows must cover own rows + margin, zones of a slab must not overlap."""
import importlib
from types import SimpleNamespace

import numpy as np
import pytest

slabs = importlib.import_module("sph-erosion_b200.slabs")


def test_ring_links():
    assert slabs.ring_links(0, 1) == (None, None, False, False)
    assert slabs.ring_links(0, 2) == (None, 1, False, False) and slabs.ring_links(1, 2) == (0, None, False, False)
    # 3+ slabs close into a ring: first and last are linked by wrap links (collisionS sends x == -len to +len)
    assert slabs.ring_links(0, 3) == (2, 1, True, False)
    assert slabs.ring_links(1, 3) == (0, 2, False, False)
    assert slabs.ring_links(2, 3) == (1, 0, False, True)
    for world in (3, 5, 8):
        for r in range(world):
            l, rt, wl, wr = slabs.ring_links(r, world)
            assert slabs.ring_links(l, world)[1] == r and slabs.ring_links(rt, world)[0] == r
            assert wl == (r == 0) and wr == (r == world - 1)


def test_channel_block_layouts():
    for layout in ("tiled", "contiguous"):
        n, world = 8, 3
        parts = [slabs.channel_block(n, world, r, False, layout=layout) for r in range(world)]
        ids = np.concatenate([p[1] for p in parts])
        assert np.array_equal(ids, np.arange(world * n ** 3)), "global ids = lattice index, no gaps"
        box = parts[0][2]
        assert box == (0.02 * n * world, 0.02 * n, 0.02 * n)
        x = np.concatenate([p[0][:, 0] for p in parts])
        assert x.min() >= -box[0] and x.max() < box[0]
        bounds = parts[0][3]
        assert len(bounds) == world - 1 and all(a < b for a, b in zip(bounds, bounds[1:]))
        for r, p in enumerate(parts):   # every block lies inside its own slab
            lo = -np.inf if r == 0 else bounds[r - 1]
            hi = np.inf if r == world - 1 else bounds[r]
            assert (p[0][:, 0] > lo).all() and (p[0][:, 0] < hi).all(), layout
    # tiled: every block is centred in its compartment
    p = slabs.channel_block(10, 4, 2, False, layout="tiled")[0][:, 0]
    L = 0.2
    centre = -4 * L + 2 * 2 * L + L
    assert abs(0.5 * (p.min() + p.max()) - centre) < 1e-5


class StubGrid:
    def __init__(self, rows, cols):
        self.rows, self.cols, self.window = rows, cols, None

    def accumulators(self):
        return 0, 0, self.rows * self.cols

    def shape(self):
        return self.rows, self.cols

    def set_window(self, a, b):
        self.window = (a, b)


def test_terrain_window_zones(monkeypatch):
    monkeypatch.setattr(slabs, "device_int32_view", lambda ptr, n, device: np.zeros(n, np.int32))
    rows, cols, world = 4096, 64, 4
    gi = SimpleNamespace(gmin=[-12.9, 0, 0], cell=0.0457446)
    colsx = slabs.partition_columns(564, world)
    cuts = slabs.terrain_row_cuts(gi, colsx, -12.8, 0.00625)
    assert len(cuts) == world - 1 and all(0 < a < b < rows for a, b in zip(cuts, cuts[1:]))
    W = slabs.terrain_margin_rows(gi.cell, 0.00625)
    assert W == int(np.ceil(3 * gi.cell / 0.00625)) + 4
    shares = [slabs.TerrainWindowShare(StubGrid(rows, cols), None, r, world, cuts, W, swap=False) for r in range(world)]
    for r, sh in enumerate(shares):
        assert sh.grid.window == sh.window
        assert sh.window[0] == max(sh.own[0] - W, 0) and sh.window[1] == min(sh.own[1] + W, rows)
        assert (sh.zone_l is None) == (r == 0) and (sh.zone_r is None) == (r == world - 1)
        if r:
            assert sh.zone_l == shares[r - 1].zone_r, "neighbours sum the same rows"
            assert sh.zone_l.stop - sh.zone_l.start == 2 * W * cols
        if sh.zone_l is not None and sh.zone_r is not None:
            assert sh.zone_l.stop <= sh.zone_r.start, "the two boundary zones of a slab do not overlap"
    assert shares[0].own[0] == 0 and shares[-1].own[1] == rows
    assert all(a.own[1] == b.own[0] for a, b in zip(shares, shares[1:])), "own rows tile the terrain"
    # slabs narrower than two zones are refused
    with pytest.raises(ValueError):
        slabs.TerrainWindowShare(StubGrid(64, cols), None, 1, 4, [16, 32, 48], W, swap=False)


def test_balanced_cuts_respect_the_minimum_width():
    hist = np.zeros(200, np.int64); hist[:20] = 1000           # everything in the first 20 columns
    free = slabs.balanced_cuts(hist, 4)
    assert free[0] < 12, "without a minimum width the first slab shrinks towards 2*HALO columns"
    cuts = slabs.balanced_cuts(hist, 4, min_width=30)
    edges = [0] + cuts + [200]
    assert all(b - a >= 30 for a, b in zip(edges, edges[1:])), cuts
    with pytest.raises(ValueError):
        slabs.balanced_cuts(hist, 4, min_width=51)


class StubGridH(StubGrid):
    """StubGrid with a heights array behind heights_device() / refresh() (numpy stands in for the device tensor)."""
    store = {}

    def __init__(self, rows, cols, fill):
        super().__init__(rows, cols)
        self.h = np.full(rows * cols, fill, np.int32)
        self.key = len(StubGridH.store) + 1
        StubGridH.store[self.key] = self.h
        self.refreshed = 0

    def heights_device(self):
        return self.key, self.h.size

    def refresh(self):
        self.refreshed += 1


def test_terrain_recut_brings_every_row_up_to_date_from_its_owner(monkeypatch):
    """TerrainWindowShare.recut_local: each replica is current only inside its window; after the re-cut every replica holds
    the OWNERS' rows everywhere, the windows sit at the new cuts and the cull maps were rebuilt."""
    class Arr(np.ndarray):
        def clone(self): return self.copy().view(Arr)
        def copy_(self, o): self[...] = o
    monkeypatch.setattr(slabs, "device_int32_view",
                        lambda ptr, n, device: StubGridH.store[ptr].view(Arr) if ptr in StubGridH.store else np.zeros(n, np.int32))
    rows, cols, world, W = 400, 8, 3, 10
    grids = [StubGridH(rows, cols, fill=-7) for _ in range(world)]          # -7 = stale everywhere
    shares = [slabs.TerrainWindowShare(grids[r], None, r, world, [100, 250], W, swap=False) for r in range(world)]
    gi = SimpleNamespace(gmin=[0.0, 0, 0], cell=0.05)
    for sh in shares:
        sh.bind_columns(gi, 0.0, 0.01)                                       # 5 terrain rows per neighbour-grid column
    for r, sh in enumerate(shares):                                          # every replica current on its window: row index * 10 + owner tag on owned rows
        h = grids[r].h.reshape(rows, cols)
        for x in range(sh.window[0], sh.window[1]):
            h[x] = 10 * x
    assert shares[0].min_columns() * 5 >= 2 * W
    new_cols = [(0, 30), (30, 60), (60, 80)]                                 # rows 150, 300
    assert shares[0].rows_of(new_cols) == [150, 300]
    slabs.TerrainWindowShare.recut_local(shares, shares[0].rows_of(new_cols))
    want = (10 * np.arange(rows, dtype=np.int32))[:, None].repeat(cols, 1)
    for r, (g, sh) in enumerate(zip(grids, shares)):
        assert np.array_equal(g.h.reshape(rows, cols), want), "replica %d" % r
        assert g.refreshed == 1 and g.window == sh.window
    assert [sh.own for sh in shares] == [(0, 150), (150, 300), (300, 400)]
    assert shares[1].window == (140, 310) and shares[1].zone_l == shares[0].zone_r
