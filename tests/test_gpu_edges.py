"""Edge cases of the hot path through the C ABI, against the oracle's ALL-PAIRS step (the restatement of the
reference's own loops, fluid_system.h:104-183) wherever particles exist: empty and tiny systems, coincident
particles (the 1e-4 branch, :438-440), everything in one cell (lists far beyond any capacity), particles on and
outside the walls (one-axis clamp, the x == -len case of collisionS :375-382), odd particle counts."""
import numpy as np
import pytest

from conftest import product
from oracle import port

pytestmark = pytest.mark.gpu
RTOL = 1e-5
FIELDS = ("density", "pos", "vel", "fpress", "fvisc", "fsurf", "normal")


def make(P, variant=None):
    m = product()
    s = m.FluidSystemSPH()
    if variant is not None:
        s.set_variant(*variant)
    q = s.params
    q.mass, q.visc, q.surf_tens, q.p0, q.k, q.h, q.len, q.dt = P.mass, P.visc, P.surf_tens, P.p0, P.k, P.h, P.len, P.dt
    q.g[0], q.g[1], q.g[2] = P.g[0], P.g[1], P.g[2]
    s.set_diagnostics(True)
    return s


def compare(s, S, tag):
    for f in FIELDS:
        got = s.download(f).astype(np.float64); want = getattr(S, f).astype(np.float64)
        scale = max(np.abs(want).max(), 1e-30)
        err = np.abs(got - want).max()
        assert err <= RTOL * scale, "%s %s: max err %.3e > %.1e * %.3e" % (tag, f, err, RTOL, scale)


def lockstep(pos, vel, P, steps, variant=None):
    """Every step starts from the oracle's state: each comparison is a one-step comparison."""
    S = port.State(pos, vel)
    s = make(P, variant)
    for k in range(steps):
        s.upload_state(S.pos, S.vel)
        s.Run()
        port.step_allpairs(P, S)
        compare(s, S, "step %d" % k)
    return s, S


def test_empty_system_steps_and_downloads():
    m = product()
    s = m.FluidSystemSPH()
    s.SetDeltaTime(0.01)
    s.Run(); s.Run()
    assert s.count() == 0
    assert s.download("pos").shape == (0, 3) and s.download("density").shape == (0,)
    s.Initialize(1000)          # and it still works afterwards
    s.Run()
    assert s.count() == 1000 and np.isfinite(s.download("pos")).all()


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_tiny_systems(n):
    rng = np.random.default_rng(n)
    pos = (rng.uniform(-0.02, 0.02, (n, 3))).astype(np.float32)
    vel = rng.normal(0, 0.3, (n, 3)).astype(np.float32)
    lockstep(pos, vel, port.default_params(dt=0.01), 3)


def test_coincident_particles_take_the_reference_direction_branch():
    """dist < 1e-4: gradPressure uses the direction (1,1,1)/sqrt(3) (fluid_system.h:438-440)."""
    pos = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [3e-5, 0.0, 0.0], [0.01, 0.01, 0.0], [0.01, 0.01, 0.0]], np.float32)
    vel = np.zeros_like(pos); vel[1] = (0.1, -0.2, 0.05)
    s, S = lockstep(pos, vel, port.default_params(dt=0.01), 2)
    assert np.abs(S.fpress).max() > 0


@pytest.mark.parametrize("variant", [(6, 3), (20, 20), (0, 0)], ids=["default", "staged", "tpp"])
def test_everything_in_one_cell(variant):
    """300 particles inside one neighbour-grid cell: every particle has 299 neighbours, every pair list is ~5x longer
    than the shared-memory stage and longer than the initial HBM rows -- spill, row growth and the direct-walk
    fallback all run, and all of them must give the reference's sums."""
    rng = np.random.default_rng(7)
    pos = rng.uniform(0.001, 0.03, (300, 3)).astype(np.float32)
    vel = rng.normal(0, 0.2, (300, 3)).astype(np.float32)
    s, S = lockstep(pos, vel, port.default_params(dt=0.0005), 8, variant)
    if variant[0] in (3, 6):
        assert s.nlist_capacity() >= 256, "the HBM rows must have grown"


def test_particles_on_and_outside_the_walls():
    """One-axis clamp with the reference's tie order; a particle EXACTLY on the -x wall goes to the +x wall
    (collisionS, fluid_system.h:375-382: `x < -len` is false, so the else branch clamps to +len)."""
    L = np.float32(0.2)
    pos = np.array([[-L, 0.0, 0.0],            # on the -x wall: lands on the +x wall
                    [L, 0.05, 0.0],            # on the +x wall: stays
                    [0.0, -L, 0.1],            # on the floor
                    [-0.25, -0.3, 0.0],        # outside on two axes: only the larger one is clamped
                    [0.3, 0.3, 0.3],           # tie on three axes: x wins (strict <)
                    [0.1, 0.19, -0.25],
                    [0.0, 0.0, 0.0]], np.float32)
    vel = np.tile(np.array([0.0, 0.02, 0.01], np.float32), (len(pos), 1))   # |v| > 0: the response divides by it
    vel[5] = (0.0, 3.0, -1.0)
    P = port.default_params(dt=0.01, g=(0.0, 0.0, 0.0))
    s, S = lockstep(pos, vel, P, 1)
    got = s.download("pos")
    assert got[0, 0] == L and S.pos[0, 0] == L, "x == -len is clamped to +len, as in the reference"
    assert np.array_equal(got.view(np.uint32), S.pos.view(np.uint32)), "clamped coordinates are exact"


@pytest.mark.parametrize("n", [999, 4097])
def test_odd_counts_random_cloud(n):
    """Odd particle counts (the kernels process particle PAIRS) in a cloud that straddles the walls."""
    rng = np.random.default_rng(n)
    pos = rng.uniform(-0.21, 0.21, (n, 3)).astype(np.float32)
    vel = rng.normal(0, 0.5, (n, 3)).astype(np.float32)
    lockstep(pos, vel, port.default_params(dt=0.005), 2)


def test_positions_into_a_device_buffer_match_the_download():
    """Renderer hand-off: sphe_write_positions_device fills a caller-owned DEVICE buffer -- the pointer a
    mapped OpenGL vertex buffer gives -- with the packed id-order positions the host download returns; a host pointer or
    a buffer that is too small is refused."""
    import torch
    m = product()
    s = m.FluidSystemSPH()
    s.Initialize(3375); s.SetDeltaTime(0.01)
    for _ in range(3):
        s.Run()
    n = s.count()
    buf = torch.zeros(3 * n + 7, dtype=torch.float32, device="cuda")
    s.write_positions_device(buf.data_ptr(), buf.numel())
    torch.cuda.synchronize()
    assert np.array_equal(buf[:3 * n].cpu().numpy().reshape(n, 3), s.download("pos"))
    assert float(buf[3 * n:].abs().sum()) == 0.0
    with pytest.raises(m.capi.SpheError):
        s.write_positions_device(buf.data_ptr(), 3 * n - 1)
    host = np.zeros(3 * n, np.float32)
    with pytest.raises(m.capi.SpheError):
        s.write_positions_device(host.ctypes.data, host.size)
