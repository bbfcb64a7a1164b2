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
