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


if __name__ == "__main__":
    default_scene()
    random_state()
    add_particles()
