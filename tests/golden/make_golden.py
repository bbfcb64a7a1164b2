"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, built in place
from /root/reference by oracle/Makefile).  Run in the authoring container only:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures pin oracle/sph_oracle.c (tests/test_oracle_golden.py) and, through it, the CUDA path.
All arrays are raw float32 bit patterns of the reference's own output.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FLOAT_FIELDS = ["pos", "vel", "acc", "density", "pressure", "fpress", "fvisc", "fgrav", "fsurf", "normal"]


def dump(sim):
    return {f: sim.field(f) for f in FLOAT_FIELDS + ["neighb"]}


def state_hash(sim):
    return ref.fnv1a64(sim.field("pos"), sim.field("vel"), sim.field("density"), sim.field("pressure"))


def mean_ke(sim, mass):
    v = sim.field("vel").astype(np.float64)
    return float((0.5 * mass * (v * v).sum(axis=1)).mean())


def default_scene():
    """BASELINE.json configs[0]: Initialize(1000), dt = 0.01 (main.cpp:183, :476-485), 100 steps."""
    s = ref.RefSim()
    s.initialize(1000)
    s.set_dt(0.01)
    out = {"pos0": s.field("pos")}
    hashes, kes, steps = [], [], []
    for step in range(1, 101):
        s.run(1)
        if step == 1:
            out.update({"s1_" + k: v for k, v in dump(s).items()})
        if step in (1, 3, 10, 20, 50, 100):
            steps.append(step); hashes.append(state_hash(s)); kes.append(mean_ke(s, 0.02))
        if step in (20, 100):
            out["s%d_pos" % step] = s.field("pos")
            out["s%d_vel" % step] = s.field("vel")
            out["s%d_density" % step] = s.field("density")
    out["hash_steps"] = np.array(steps, np.int32)
    out["hashes"] = np.array(hashes)
    out["mean_ke"] = np.array(kes, np.float64)
    np.savez_compressed(os.path.join(OUT, "default_scene.npz"), **out)
    print("default_scene", dict(zip(steps, hashes)))


def random_state():
    """Random gas in a +-0.3 cube with velocities, several parameter sets incl. dt = 0 and a
    box-crossing case; 1 step each."""
    rng = np.random.default_rng(0x5EED)
    cases = {}
    n = 1500
    for ci, (dt, length, h) in enumerate([(0.01, 0.2, 0.0457), (0.0, 0.2, 0.0457), (0.004, 0.3, 0.06)]):
        pos = rng.uniform(-length * 1.02, length * 1.02, (n, 3)).astype(np.float32)
        vel = rng.normal(0, 0.5, (n, 3)).astype(np.float32)
        s = ref.RefSim()
        s.set_len(length); s.set_h(h); s.set_dt(dt)
        s.set_params(0.021, 3.0, 0.07, 1000.0, [0.1, -9.0, 0.2])
        s.set_state(pos, vel)
        s.run(1)
        d = dump(s)
        cases.update({"c%d_in_pos" % ci: pos, "c%d_in_vel" % ci: vel,
                      "c%d_cfg" % ci: np.array([dt, length, h, 0.021, 3.0, 0.07, 1000.0, 0.1, -9.0, 0.2], np.float32)})
        cases.update({"c%d_%s" % (ci, k): v for k, v in d.items()})
    np.savez_compressed(os.path.join(OUT, "random_state.npz"), **cases)
    print("random_state ok")


def add_particles():
    """Initialize(1000) + AddParticles(125) at the same start coordinates: coincident particles hit
    the dist < 1e-4 branch of gradPressure (fluid_system.h:438-440).  Then Reset()."""
    s = ref.RefSim()
    s.set_origin(0.01, 0.02, -0.01)
    s.initialize(1000)
    s.set_dt(0.01)
    s.run(2)
    s.add_particles(125)
    out = {"added_pos": s.field("pos"), "added_vel": s.field("vel"), "added_id": s.field("id")}
    s.run(1)
    out.update({"a1_" + k: v for k, v in dump(s).items()})
    s.reset()
    out["reset_pos"] = s.field("pos")
    out["reset_id"] = s.field("id")
    np.savez_compressed(os.path.join(OUT, "add_particles.npz"), **out)
    print("add_particles ok", s.count())


def terrain_trials(rng, img, n, pen, step, lo=2.5, hi=46.5):
    """Particle moves ending near the heightfield surface (cells stay inside [1, 47]: at the border the
    reference reads heights at index -1, undefined behaviour)."""
    x = rng.uniform(lo, hi, n); z = rng.uniform(lo, hi, n)
    hloc = img[np.floor(x).astype(int), np.floor(z).astype(int)].astype(np.float32)
    vdir = rng.normal(0, 1, (n, 3)); vdir[:, 1] = -np.abs(vdir[:, 1]) * 2 - 0.2
    vdir /= np.linalg.norm(vdir, axis=1)[:, None]
    pn = np.stack([x, hloc - rng.uniform(0, pen, n) + rng.uniform(-3, 3, n), z], 1)
    pc = pn - vdir * np.minimum(rng.uniform(0.01, step, n), 1.4)[:, None]
    vn = vdir * rng.uniform(0.1, 5, n)[:, None]
    return pc.astype(np.float32), pn.astype(np.float32), vn.astype(np.float32)


def terrain():
    """Grid::collision (grid.h:462-805) and UpdateGrid/genIndices (grid.h:118-176) of the unmodified
    reference on lena_gray.png, Grid(50, 255, 50) as in main.cpp:100-105.  Only the top-left 64 x 64
    texels matter for these inputs; they are stored so the test needs no PNG decoder."""
    from PIL import Image
    img = np.array(Image.open("/root/reference/Erosion/lena_gray.png"))
    assert img.shape == (512, 512) and img.dtype == np.uint8
    g = ref.RefGrid(50, 255, 50); g.load_heightfield(img); g.update(50, 255, 50)
    rng = np.random.default_rng(0x7E44A1)
    out = {"hf64": img[:64, :64].copy(), "dims": np.array([50, 255, 50], np.int32)}
    pcs, pns, vns = [], [], []
    for pen, step in [(0.02, 0.05), (0.5, 0.5), (3, 1.5), (8, 3.0), (0.1, 1.2)]:
        pc, pn, vn = terrain_trials(rng, img, 1600, pen, step)
        pcs.append(pc); pns.append(pn); vns.append(vn)
    # degenerate inputs: zero velocity (NaN direction), vertical drop onto a vertex, a move along a cell edge
    pcs.append(np.array([[10.5, 200.0, 10.5], [12.0, 200.0, 12.0], [20.0, float(img[20, 20]) + 0.4, 20.25]], np.float32))
    pns.append(np.array([[10.5, 10.0, 10.5], [12.0, float(img[12, 12]) - 0.5, 12.0], [20.0, float(img[20, 20]) - 0.3, 20.75]], np.float32))
    vns.append(np.array([[0, 0, 0], [0, -3, 0], [0, -1, 1]], np.float32))
    pc, pn, vn = np.concatenate(pcs), np.concatenate(pns), np.concatenate(vns)
    hit, cp, nrm = g.collision(pc, pn, vn)
    out.update(pc=pc, pn=pn, vn=vn, hit=hit, cp=cp, nrm=nrm)
    out["surface"] = g.surface(); out["indices"] = g.indices()
    out["heights_probe"] = np.array([g.height_at(x, z) for x, z in [(0, 0), (5, 7), (49, 49), (63, 1)]], np.int32)
    out["voxel_probe"] = np.array([[g.voxel_type(x, int(img[x, z]), z), g.voxel_type(x, int(img[x, z]) - 1, z)] for x, z in [(3, 4), (20, 31)]], np.int32)
    np.savez_compressed(os.path.join(OUT, "terrain.npz"), **out)
    print("terrain ok: %d trials, %d hits (%.1f%%), surface %d floats, %d indices" % (len(hit), hit.sum(), 100 * hit.mean(), out["surface"].size, out["indices"].size))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "terrain":
        terrain(); sys.exit(0)
    default_scene()
    random_state()
    add_particles()
    terrain()
