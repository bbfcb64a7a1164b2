"""oracle/sph_oracle.c against the UNMODIFIED reference compiled in place (oracle/_ref).
Skipped when the prebuilt library is absent (it is git-ignored; `make -C oracle ref` builds it
where /root/reference exists)."""
import numpy as np
import pytest

from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
FLOAT_FIELDS = ["pos", "vel", "acc", "density", "pressure", "fpress", "fvisc", "fgrav", "fsurf", "normal"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_particle_struct_size():
    assert ref._load(False).ref_sizeof_particle() == 112


@pytest.mark.parametrize("n,length", [(1000, 0.2), (3375, 0.3)])
def test_lattice_scenes_bit_exact(n, length):
    r = ref.RefSim(); r.set_len(length); r.initialize(n); r.set_dt(0.01)
    P = port.default_params(dt=0.01, len=length)
    A = port.State(port.lattice(n)); B = port.State(port.lattice(n))
    G = port.grid_for_box(P, [-length - 0.05] * 3, [length + 0.05] * 3)
    for step in range(6):
        r.run(1); port.step_allpairs(P, A); port.step_grid(P, G, B)
        for f in FLOAT_FIELDS:
            assert np.array_equal(bits(r.field(f)), bits(getattr(A, f))), (step, f)
            assert np.array_equal(bits(r.field(f)), bits(getattr(B, f))), (step, f, "grid")
        assert np.array_equal(r.field("neighb"), A.neighb)


def test_serial_and_openmp_reference_agree():
    if not ref.available(omp=True):
        pytest.skip("omp build absent")
    a = ref.RefSim(omp=False); b = ref.RefSim(omp=True)
    for s in (a, b):
        s.initialize(1000); s.set_dt(0.01); s.run(5)
    for f in FLOAT_FIELDS:
        assert np.array_equal(bits(a.field(f)), bits(b.field(f))), f


def test_random_gas_with_params():
    rng = np.random.default_rng(11)
    n = 1200
    pos = rng.uniform(-0.21, 0.21, (n, 3)).astype(np.float32)
    vel = rng.normal(0, 1.0, (n, 3)).astype(np.float32)
    r = ref.RefSim(); r.set_dt(0.005); r.set_params(0.018, 2.5, 0.05, 990.0, [0.0, -9.82, 0.3]); r.set_state(pos, vel)
    P = port.default_params(dt=0.005, mass=0.018, visc=2.5, surf_tens=0.05, p0=990.0, g=(0.0, -9.82, 0.3))
    S = port.State(pos, vel)
    G = port.grid_for_box(P, [-0.3] * 3, [0.3] * 3)
    for step in range(4):
        r.run(1); port.step_grid(P, G, S)
        for f in FLOAT_FIELDS:
            assert np.array_equal(bits(r.field(f)), bits(getattr(S, f))), (step, f)
