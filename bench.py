#!/usr/bin/env python
"""bench.py -- particle-updates/s of the SPH-Erosion per-step hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|small] [--impl reference]

One "step" = one FluidSystemSPH::Run over the synthetic scene: hash -> count/scan -> scatter ->
rank+reorder -> density -> force+integrate+collide (+terrain/erosion when the workload has terrain).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what every key means.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sph_particle_updates_per_sec"
UNIT = "particle-updates/s"
SPACING = 0.025
# Algorithmic HBM bytes per particle per launch: SURVEY.md section 8(d)'s table, the official figure behind roofline.achieved
# (hash 24; sort 52 + cell-start 6 -> here the counting sort's scan + scatter; reorder 68; density 20; force+integrate 64;
# whole step ~234).  The terrain stage is not in that table (no reference erosion): builder's figure.
ALGO_BYTES = {"hash": 24, "scan": 6, "scatter": 52, "reorder": 68, "density": 20, "force": 64, "terrain": 92}
STEP_BYTES = 234
# What THIS implementation's arrays make compulsory per launch, reported beside it as "layout":
# density also writes posC + m/rho and reads the cell ids, force also reads rho + the list segment counts, reorder also
# moves ids / sediment / cell ids.  Neighbour-list traffic (written by one pass, read by the other) is NOT in either figure.
LAYOUT_BYTES = {"hash": 20, "scan": 0, "scatter": 16, "reorder": 92, "density": 44, "force": 72, "terrain": 92}


# ----------------------------------------------------------------------------- scenes
def scaled_dam_break(n_axis, jitter=False, nx_mult=1):
    """box half-extent L = 0.02*n, lattice spacing 0.025, block corner at
    (-L, -L/4, -3L/4); n_axis = 10 reproduces the reference default scene exactly.  nx_mult stretches
    the block (and the box) along x for weak scaling over x-slabs."""
    L = np.float32(0.02 * n_axis * nx_mult)
    Lb = 0.02 * n_axis
    ix = np.arange(n_axis * nx_mult)
    i = np.arange(n_axis)
    x = (-float(L) + ix * SPACING).astype(np.float32)
    y = (-Lb / 4 + i * SPACING).astype(np.float32)
    z = (-0.75 * Lb + i * SPACING).astype(np.float32)
    pos = np.empty((x.size, y.size, z.size, 3), np.float32)
    pos[..., 0] = x[:, None, None]; pos[..., 1] = y[None, :, None]; pos[..., 2] = z[None, None, :]
    pos = pos.reshape(-1, 3)
    if jitter:
        rng = np.random.default_rng(0x5EED)
        pos += rng.uniform(-0.2 * SPACING, 0.2 * SPACING, pos.shape).astype(np.float32)
    return pos, float(L)


def scene_gravity(n_axis, unscaled=False):
    """Dynamic similarity with the reference default scene (n_axis = 10): the soft equation of state
    (k = 3, fluid_system.h:467) lets a column of height H compress by ~ g*H/k, so a scene scaled by
    n/10 with unscaled g free-falls 1.5 box units and compresses to ~40x rest density (measured:
    180 neighbours/particle at 1M particles) -- neither the reference's regime nor a stationary
    workload.  Scaling g by 10/n keeps g*H/k, the velocities and the neighbour counts of the reference
    scene (15-33 per particle).  g is a UI-mutable parameter (GetGrav, :281-284)."""
    return -9.82 if unscaled else -9.82 * 10.0 / n_axis


WORKLOADS = {
    # name: (n_axis, jitter, terrain, description)
    "small": (32, False, False, "32^3 = 32,768-particle dam break (smoke-sized)"),
    "c2": (100, False, False, "BASELINE configs[1]: 1M-particle dam break in a box, no terrain (n=100 -> 1,000,000)"),
    "c3": (160, True, True, "BASELINE configs[2]: 4.096M-particle SPH erosion (n=160) over a 1024x1024 heightmap terrain, sediment pickup/deposit enabled"),
    "c3-noterrain": (160, True, False, "the c3 particle scene without the terrain stage (4.096M-particle dam break, n=160)"),
}


def synthetic_heightmap(size=1024, seed=0x7E44A1):
    """8-bit heightmap in the style of the reference's photographs (lena_gray.png: min 34, max 246, mean
    125, rough at the pixel scale): five octaves of bilinear value noise plus pixel noise.  The PNGs are
    reference assets and do not travel to the GPU box, so the scene generates its own."""
    rng = np.random.default_rng(seed)
    h = np.zeros((size, size), np.float64)
    amp, total = 1.0, 0.0
    for cells in (4, 8, 16, 64, 256):
        g = rng.uniform(0, 1, (cells + 1, cells + 1))
        t = np.linspace(0, cells, size, endpoint=False)
        i = t.astype(int); f = t - i
        f = f * f * (3 - 2 * f)
        row = g[i][:, i] * (1 - f)[None, :] + g[i][:, i + 1] * f[None, :]
        row1 = g[i + 1][:, i] * (1 - f)[None, :] + g[i + 1][:, i + 1] * f[None, :]
        h += amp * (row * (1 - f)[:, None] + row1 * f[:, None])
        total += amp; amp *= 0.5
    h = h / total + rng.normal(0, 0.01, h.shape)
    h = (h - h.min()) / (h.max() - h.min())
    return np.clip(np.rint(34 + h * (246 - 34)), 0, 255).astype(np.uint8)


def heightmap_1024():
    """the reference's Erosion/lena_gray.png mirrored 2 x 2 to 1024 x 1024, exact bytes (fixture written
    by tests/golden/make_heightmap.py from the reference asset); the synthetic map only if the fixture is missing."""
    p = os.path.join(ROOT, "tests", "golden", "lena_gray_512.npz")
    if os.path.exists(p):
        a = np.load(p)["heights_u8"]
        top = np.concatenate([a, a[:, ::-1]], axis=1)
        return np.ascontiguousarray(np.concatenate([top, top[::-1, :]], axis=0)), \
            "Erosion/lena_gray.png (512 x 512, 8-bit) mirrored 2 x 2 to 1024 x 1024, exact bytes (tests/golden/lena_gray_512.npz)"
    return synthetic_heightmap(), "synthetic 1024x1024 8-bit value noise (lena_gray statistics), seed 0x7E44A1 -- lena fixture missing"


def attach_terrain(pkg, L, n_axis, nx_mult=1):
    """1024 x 1024 terrain under the whole box floor: cell = 2L/1024 world units (uniform scale), heights
    0..255 levels scaled to 0..0.25*block height so the relief is a fraction of the fluid depth, terrain
    top just under the block so the fluid lands on it during the settle phase.  nx_mult > 1 (multi-GPU weak
    scaling): the heightmap is repeated nx_mult times along x under the (nx_mult*L, L, L) channel."""
    base, what = heightmap_1024()
    img = np.tile(base, (nx_mult, 1))
    cell = 2.0 * L / 1024.0
    relief = 0.1 * L                                    # world units between the lowest and highest vertex
    heights = img.astype(np.float32) * np.float32(relief / 255.0 / cell)   # in cells
    g = pkg.Grid(1024 * nx_mult, 255, 1024)
    g.set_heights(heights)
    top = float(heights.max()) * cell
    origin = (-L * nx_mult, -L / 4 - 0.5 * SPACING - top, -L)     # highest vertex half a lattice spacing under the block
    g.set_transform(origin, cell)
    e = g.erosion
    e.enabled = 1; e.Kc = 2.0; e.Ke = 0.3; e.Kd = 0.3; e.hmin = 0.0; e.max_pickup = 0.25
    return g, {"heightmap": what + (", repeated %d x along x" % nx_mult if nx_mult > 1 else ""),
               "terrain_cell": cell, "terrain_relief": relief, "terrain_origin": list(origin),
               "erosion": {"Kc": 2.0, "Ke": 0.3, "Kd": 0.3, "hmin": 0.0, "max_pickup": 0.25}}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  NVML in a thread (a sample every ~2 ms: the
    default timed region is only ~40 ms long); falls back to `nvidia-smi -lms` when pynvml is unavailable."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index
        self.nvml = None
        self.samples = []      # (sm_mhz, reasons bitmask)
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    phys = self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)),
                                     int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons")
                                     else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [a for a, _ in self.samples]
            mask = 0
            for _, m in self.samples:
                mask |= m
            reasons = sorted(k for k, bit in names.items() if mask & bit)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, one sample per ~2 ms during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


# ----------------------------------------------------------------------------- helpers
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class near_gpu:
    """Context manager: run the enclosed host code on the CPUs NVML names as local to GPU `index` (same NUMA node / PCIe root),
    so pinned buffers allocated and first touched inside land in memory next to that GPU.  Restores the affinity on exit
    (the CPU baselines use every core).  A no-op when NVML or sched_setaffinity is unavailable.  .cpus = what was set."""

    def __init__(self, index):
        self.index, self.old, self.cpus = index, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    phys = self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            old = os.sched_getaffinity(0)
            cpus &= old
            if cpus and cpus != old:
                os.sched_setaffinity(0, cpus)
                self.old, self.cpus = old, sorted(cpus)
        except Exception:
            self.old = None
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except OSError:
                pass
        return False


def ncu_traffic(kernel, workload="c2"):
    """dram read+write bytes per launch of `kernel` on `workload` from the committed ncu summaries
    (profiles/ncu_traffic.json), or None when that kernel/workload pair has not been captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload, {}).get(kernel)
        except Exception:
            return None
    return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------- CPU arms
def cpu_port_baseline(pos, L, gy=-9.82, budget_s=20.0, terrain=None, vel=None):
    """Times oracle/sph_oracle.c so_step_grid (OpenMP, all host cores) on a bounded sample of the
    same scene.  kind = "port": the cell-grid restatement, bit-identical to the reference's sums.
    terrain = (heights_in_cells, origin, cell, erosion dict): the terrain stage of the oracle runs too."""
    from oracle import port
    n = pos.shape[0]
    sample_n = n
    # ~1e6 updates/s on 8 cores: keep the sample around the budget
    est = n / (1.2e5 * max(port.omp_threads(), 1))
    if est > budget_s:
        sample_n = int(n * budget_s / est)
    # a contiguous x-slab of the block keeps the neighbour statistics of the full scene
    order = np.argsort(pos[:, 0], kind="stable")[:sample_n]
    sp = np.ascontiguousarray(pos[order])
    P = port.default_params(dt=0.01, len=L, g=(0.0, gy, 0.0))
    S = port.State(sp, None if vel is None else vel[order])
    G = port.grid_for_box(P, [-L - 0.1] * 3, [L + 0.1] * 3)
    steps = 2
    if terrain is None:
        port.step_grid(P, G, S)  # warm (page faults, thread pool)
        t = time.perf_counter()
        port.step_grid(P, G, S, steps)
        what = "oracle so_step_grid (OpenMP cell grid, sums bit-identical to the reference)"
    else:
        heights, origin, cell, ero = terrain
        T = port.Terrain(heights, (heights.shape[0], 255, heights.shape[1]))
        E = port.erosion_params(enabled=True, origin=origin, scale=cell, **ero)
        sed = np.zeros(S.n, np.int32)
        port.step_grid_terrain(P, G, S, T, E, sed)
        t = time.perf_counter()
        for _ in range(steps):
            port.step_grid_terrain(P, G, S, T, E, sed)
        what = "oracle so_forces_grid + so_terrain_stage + so_box (OpenMP cell grid + terrain contact/erosion restatement)"
    dt = (time.perf_counter() - t) / steps
    return {"value": sample_n / dt, "unit": UNIT, "cores": port.omp_threads(), "kind": "port",
            "sample": "%d of %d particles (x-slab of the same scene), %d steps of %s, %.2f s/step" % (sample_n, n, steps, what, dt)}


def run_reference_arm(args):
    """--impl reference: the UNMODIFIED reference step (oracle/_ref/libsphref_omp.so, built in place
    from the reference sources) on the host cores.  It is O(N^2), so each step runs a bounded sample
    of the workload's scene."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import ref
    n_axis, jitter, _, desc = WORKLOADS[args.workload]
    line = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic"}
    if not ref.available(omp=True):
        os.environ["OMP_NUM_THREADS"] = os.environ.get("SPHE_REF_THREADS", str(os.cpu_count() or 1))
        # the oracle always exists: fall back to the C port of the same algorithm
        pos, L = scaled_dam_break(n_axis, jitter)
        cb = cpu_port_baseline(pos, L, scene_gravity(n_axis, args.gravity_unscaled))
        line.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "config": {"workload": args.workload, "description": desc, "note": "oracle/_ref absent: timed the C port"},
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(line)
        return
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank); the OpenMP runtime reads the
    # variable when the reference library is loaded, which happens below
    os.environ["OMP_NUM_THREADS"] = os.environ.get("SPHE_REF_THREADS", str(os.cpu_count() or 1))
    # same parameters as the GPU arm, g included (scene_gravity of the WORKLOAD's n_axis); only the particle count is bounded
    gy_ref = scene_gravity(n_axis, args.gravity_unscaled)

    def make(sample_axis):
        pos, L = scaled_dam_break(sample_axis, jitter)
        sim = ref.RefSim(omp=True)
        sim.set_len(L); sim.set_dt(0.01); sim.set_params(0.02, 3.5, 0.0728, 998.29, [0.0, gy_ref, 0.0])
        sim.set_state(pos, np.zeros_like(pos))
        return sim, pos

    sample_axis = 22
    sim, pos = make(sample_axis)
    t = time.perf_counter(); sim.run(1); t1 = time.perf_counter() - t
    if t1 * (args.steps + args.warmup) > 150.0:
        sample_axis = 16
        sim, pos = make(sample_axis)
    sim.run(args.warmup)
    t = time.perf_counter(); sim.run(args.steps); dt = (time.perf_counter() - t) / max(args.steps, 1)
    n = pos.shape[0]
    full_n = n_axis ** 3
    cores = ref._load(True).ref_omp_max_threads()
    extrap = (n / dt) * n / full_n
    cb = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "reference",
          "sample": "%d^3 = %d-particle block of the same scaled dam break (same h, dt, g = %.3f as the GPU arm; the reference has no terrain "
                    "stage: Grid::collision is commented out at its only call site, fluid_system.h:335-340); the reference is all-pairs O(N^2), "
                    "so at the workload's %d particles it extrapolates to %.3g updates/s" % (sample_axis, n, gy_ref, full_n, extrap)}
    line.update({"value": n / dt, "ms_per_step": dt * 1e3, "cpu_baseline": cb,
                 "same_config": False,
                 "extrapolated_full_n_value": extrap,
                 "extrapolation": "value x sample_particles / workload_particles: one reference step costs 3 all-pairs passes, time ~ N^2 (SURVEY.md section 6)",
                 "config": {"workload": args.workload, "description": desc, "particles": full_n, "reference_sample_particles": n, "gravity_y": gy_ref},
                 "e2e": {"value": n / dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def oracle_like(sim, port):
    gi = sim.grid_info()
    G = port.Grid(); G.gmin[:] = list(gi.gmin); G.cell = gi.cell; G.dim[:] = list(gi.dim)
    return G


def parity_gate(sim, grid, L, gy, tinfo):
    """In-bench parity at FULL size: one more step from the state the timed region left, the same step by the oracle
    (oracle/sph_oracle.c + terrain_oracle.c: the checker, never the thing measured) on the same input, all particles.
    Bars as in tests/: binning bit-exact; density, positions, velocities within RTOL = 1e-5 of the field's scale (20x
    that with a terrain: the contact response amplifies 1e-5 velocity differences, tests/test_gpu_terrain.py); terrain
    heights within 8 fixed-point units (1/4096 height unit) on every vertex, carried sediment total within the same."""
    from oracle import port
    t0 = time.perf_counter()
    pos = sim.download("pos"); vel = sim.download("vel")
    sed = np.rint(sim.download("sediment").astype(np.float64) * 4096.0).astype(np.int32) if grid is not None else None
    h0 = grid.heights() if grid is not None else None
    sim.Run(grid)
    P = port.default_params(dt=0.01, len=L, g=(0.0, gy, 0.0))
    G = oracle_like(sim, port)
    cell_of, order, cell_start = port.bin_particles(G, pos)
    out = {"particles": int(pos.shape[0]), "oracle": "so_forces_grid + so_terrain_stage + so_box, all particles" if grid is not None else "so_step_grid, all particles"}
    out["binning_bit_exact"] = bool(np.array_equal(sim.debug_sorted_order(), order) and np.array_equal(sim.debug_cell_start(), cell_start))
    S = port.State(pos, vel)
    if grid is not None:
        T = port.Terrain(h0, (h0.shape[0], 255, h0.shape[1]))
        E = port.erosion_params(enabled=True, origin=tinfo["terrain_origin"], scale=tinfo["terrain_cell"], **tinfo["erosion"])
        hit = port.step_grid_terrain(P, G, S, T, E, sed)
        out["oracle_contacts"] = int(hit.sum())
        dh = np.abs(grid.heights_fx().astype(np.int64) - T.hfx.astype(np.int64))
        out["height_max_diff_fx"] = int(dh.max()); out["height_vertices_differing"] = int((dh > 0).sum())
        out["sediment_total_diff_fx"] = int(abs(int(sim.sediment_total_fx()) - int(sed.astype(np.int64).sum())))
    else:
        port.step_grid(P, G, S)
    tol = 20e-5 if grid is not None else 1e-5
    ok = out["binning_bit_exact"]
    for name, want, t in (("density", S.density, 1e-5), ("pos", S.pos, tol), ("vel", S.vel, tol)):
        got = sim.download(name).astype(np.float64); want = want.astype(np.float64)
        err = float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
        out[name + "_max_err_over_scale"] = err
        ok = ok and err <= t
    rel = np.abs(sim.download("density").astype(np.float64) - S.density) / np.maximum(S.density, 1e-30)
    out["density_worst_particle_rel_err"] = float(rel.max())
    if grid is not None:
        ok = ok and out["height_max_diff_fx"] <= 8
    out["tolerance"] = {"density": 1e-5, "pos_vel": tol, "heights_fx": 8 if grid is not None else None}
    out["seconds"] = round(time.perf_counter() - t0, 2)
    return bool(ok), out


def reference_gravity_run(args, pkg, local, pos, L, n_axis, terrain, budget_s=30.0):
    """The same scene with the reference's default g = -9.82 (fluid_system.h:460) instead of the scaled one: a second,
    labelled value.  Bounded: at most min(steps, 20) timed steps, and the settle phase stops when the budget is spent."""
    import torch
    sim = pkg.FluidSystemSPH(device=local)
    sim.params.len = L; sim.params.g[1] = -9.82; sim.SetDeltaTime(0.01)
    sim.set_variant(args.density_variant, args.force_variant)
    sim.set_stream(torch.cuda.current_stream().cuda_stream)
    sim.upload_state(pos, np.zeros_like(pos))
    grid = attach_terrain(pkg, L, n_axis)[0] if terrain else None
    t = time.perf_counter(); settled = 0
    while grid is not None and settled < args.settle and time.perf_counter() - t < budget_s:
        for _ in range(10):
            sim.Run(grid)
        settled += 10; torch.cuda.synchronize()
    for _ in range(args.warmup):
        sim.Run(grid)
    k = max(3, min(args.steps, 20))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k):
        sim.Run(grid)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    n = pos.shape[0]
    nb = int(sim.debug_neighbours_total()) / n if n <= 4200000 else None
    return {"gravity_y": -9.82, "value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": k, "settle_steps": settled,
            "mean_neighbours_after_run": nb, "neighbour_list_rows": sim.nlist_capacity(),
            "note": "reference default g on a scene %d x the reference's: the soft equation of state (k = 3) lets the column compress far beyond the "
                    "reference's regime (bench.scene_gravity); reported for completeness, the headline uses the dynamically similar g" % (n_axis // 10)}


def run_gpu_arm(args):
    import torch
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("sph-erosion_b200")
    n_axis, jitter, terrain, desc = WORKLOADS[args.workload]

    if world > 1:
        from importlib import import_module
        slabs = import_module("sph-erosion_b200.slabs")
        return slabs.bench_multi(args, pkg, n_axis, jitter, desc, METRIC, UNIT, terrain)

    pos, L = scaled_dam_break(n_axis, jitter)
    n = pos.shape[0]
    gy = scene_gravity(n_axis, args.gravity_unscaled)
    sim = pkg.FluidSystemSPH(device=local)
    sim.params.len = L
    sim.params.g[1] = gy
    sim.SetDeltaTime(0.01)
    sim.set_variant(args.density_variant, args.force_variant)
    # ONE timing method for N = 1 and N > 1 (slabs.bench_multi): the step runs on torch's current stream and ONE pair of CUDA
    # events brackets all K steps; no flush between steps -- the working set of every workload exceeds the 126 MB L2
    sim.set_stream(torch.cuda.current_stream().cuda_stream)
    sim.upload_state(pos, np.zeros_like(pos))
    if args.l2_flush_mib > 0:
        sim.set_l2_flush(args.l2_flush_mib << 20)
    grid, tinfo = (attach_terrain(pkg, L, n_axis) if terrain else (None, {}))
    if grid is not None:
        # untimed settle phase: let the block land on the terrain so the timed steps exercise contacts
        for _ in range(args.settle):
            sim.Run(grid)
        tinfo["settle_steps"] = args.settle
        tot0 = grid.total_fx() + sim.sediment_total_fx()

    # The headline loop carries NO per-kernel instrumentation: an event pair around every launch costs ~9 % of a
    # 1M-particle step (scripts/event_overhead.py).  The per-kernel times of the roofline come from a second pass over
    # the SAME step window, replayed from a checkpoint taken before the warm-up (sphe_save_state; resumed runs are bit-exact).
    import tempfile
    ck = os.path.join(tempfile.mkdtemp(prefix="sphe_bench_"), "window.sphe")
    sim.save_state(ck, grid)
    for _ in range(args.warmup):
        sim.Run(grid)
    if grid is not None:
        grid.contacts(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.kernel_timing(False)   # also zeroes the launch counter
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        sim.Run(grid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    _, launches = sim.kernel_times()
    clocks = sampler.stop()
    contacts = grid.contacts() if grid is not None else 0
    sed_end = sim.sediment_total_fx() if grid is not None else 0
    total_end = (grid.total_fx() + sed_end) if grid is not None else 0
    state_end = (sim.download("pos"), sim.download("vel"), grid.heights() if grid is not None else None)
    nbr_end = int(sim.debug_neighbours_total()) if n <= 4200000 else None
    rows_end, smem_end, ovf_end = sim.nlist_capacity(), sim.nlist_smem_entries(), sim.nlist_overflowed()
    gate_ok, gate = (None, None) if args.no_parity_gate else parity_gate(sim, grid, L, gy, tinfo)
    # second pass, instrumented
    sim.load_state(ck, grid)
    os.remove(ck); os.rmdir(os.path.dirname(ck))
    sim.timed_steps(args.warmup, grid=grid, per_kernel=False)
    ms_instrumented, per_kernel, _ = sim.timed_steps(args.steps, grid=grid, per_kernel=True)
    if grid is not None:
        tinfo["terrain_contacts_per_step"] = contacts / args.steps
        tinfo["contact_cull_survivors_last_step"] = dict(zip(("same_cell", "one_axis", "both_axes"), sim.terrain_survivors()))
        tinfo["sediment_in_flight_fx"] = sed_end
        tinfo["conservation_exact"] = bool(total_end == tot0)
        tinfo["replayed_window_bit_identical"] = bool(np.array_equal(state_end[0], sim.download("pos")) and np.array_equal(state_end[2], grid.heights()))
    ms_step = ms / args.steps
    value = n / (ms_step * 1e-3)

    # end to end: host (pinned) buffers in, host buffers out, every step, through sphe_step_host
    e2e_steps = max(3, min(args.steps, 20))
    with near_gpu(local) as numa:      # pinned buffers on the NUMA node next to the GPU
        hp = torch.from_numpy(pos).pin_memory(); hv = torch.zeros_like(hp).pin_memory()
        op = torch.zeros_like(hp).pin_memory(); ov = torch.zeros_like(hp).pin_memory()
        orho = torch.zeros(n, dtype=torch.float32).pin_memory()
    sim2 = pkg.FluidSystemSPH(device=local)
    sim2.params.len = L; sim2.params.g[1] = gy; sim2.SetDeltaTime(0.01)
    sim2.set_variant(args.density_variant, args.force_variant)
    if grid is not None:
        # start the end-to-end loop from the settled state so it, too, runs with terrain contacts
        hp.copy_(torch.from_numpy(sim.download("pos"))); hv.copy_(torch.from_numpy(sim.download("vel")))
    for _ in range(3):
        sim2.step_host_ptr(n, hp.data_ptr(), hv.data_ptr(), op.data_ptr(), ov.data_ptr(), orho.data_ptr(), grid=grid)
    t = time.perf_counter()
    for _ in range(e2e_steps):
        sim2.step_host_ptr(n, hp.data_ptr(), hv.data_ptr(), op.data_ptr(), ov.data_ptr(), orho.data_ptr(), grid=grid)
        hp, op = op, hp
        hv, ov = ov, hv
    e2e_dt = (time.perf_counter() - t) / e2e_steps
    e2e = {"value": n / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 28 * n,
           "ms_per_step": e2e_dt * 1e3, "steps": e2e_steps,
           "api": "sphe_step_host (pinned host pos/vel in, pos/vel/density out, id order)",
           "pinned_buffers": ("allocated from the %d CPUs NVML lists as local to the GPU" % len(numa.cpus)) if numa.cpus else "default placement"}

    peak, peak_src = measured_peak()
    dom = max(("density", "force", "terrain"), key=lambda k: per_kernel[k])
    t_dom = per_kernel[dom] / args.steps * 1e-3
    achieved = ALGO_BYTES[dom] * n / t_dom / 1e9
    frac_of = lambda table: {k: (table[k] * n / (per_kernel[k] / args.steps * 1e-3) / 1e9 / peak)
                             for k in table if per_kernel.get(k, 0) > 0 and table[k] > 0}
    roofline = {"bound": "hbm", "kernel": "k_%s" % dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(dom, args.workload), "peak_source": peak_src,
                "algorithmic_bytes_per_particle": ALGO_BYTES[dom],
                "algorithmic_bytes_source": "SURVEY.md section 8(d) table (terrain: builder's figure, the table has none)",
                "whole_step": {"bytes_per_particle": STEP_BYTES + (ALGO_BYTES["terrain"] if grid is not None else 0),
                               "GB/s": (STEP_BYTES + (ALGO_BYTES["terrain"] if grid is not None else 0)) * n / (ms_step * 1e-3) / 1e9,
                               "frac": (STEP_BYTES + (ALGO_BYTES["terrain"] if grid is not None else 0)) * n / (ms_step * 1e-3) / 1e9 / peak},
                "note": "neighbour passes are fp32-issue/LSU bound, not HBM bound (DESIGN.md); "
                        "frac is reported against HBM as the contract asks",
                "per_kernel_timing": "second pass over the same %d-step window (replayed from a checkpoint) with a CUDA-event pair around every "
                                     "launch; that instrumentation makes the step %.1f %% slower than the headline loop, which has none"
                                     % (args.steps, 100.0 * (ms_instrumented / ms - 1.0)),
                "per_kernel_ms_per_step": {k: v / args.steps for k, v in per_kernel.items()},
                "per_kernel_hbm_frac": frac_of(ALGO_BYTES),
                "per_kernel_hbm_frac_note": "SURVEY 8(d) bytes / live time / peak; a value above 1 (scatter) means this implementation moves "
                                            "fewer bytes than the SURVEY's per-particle figure assumes, not that HBM ran above its peak",
                "per_kernel_traffic": {k: ncu_traffic(k, args.workload) for k in per_kernel if ncu_traffic(k, args.workload)},
                "layout": {"what": "the same with the bytes this implementation's arrays make compulsory (DESIGN.md section 3), labelled, not the official figure",
                           "bytes_per_particle": LAYOUT_BYTES, "per_kernel_hbm_frac": frac_of(LAYOUT_BYTES)}}
    # secondary figure: the bytes the neighbour passes GATHER -- candidates x 16 B per pass -- served by
    # L1/L2, never to be read as DRAM traffic.  Candidates of a particle = population of its 27 cells, from the cell table.
    if n <= 4200000:
        gi = sim.grid_info()
        occ = np.diff(sim.debug_cell_start().astype(np.int64)).reshape(int(gi.dim[0]), int(gi.dim[1]), int(gi.dim[2]))
        pad = np.pad(occ, 1)
        box = sum(pad[1 + dx:1 + dx + occ.shape[0], 1 + dy:1 + dy + occ.shape[1], 1 + dz:1 + dz + occ.shape[2]]
                  for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1))
        cand = float((occ * box).sum()) / n
        roofline["gather"] = {"what": "neighbour-candidate reads of the %s pass, 16 B each, served from L1/L2 -- NOT DRAM traffic" % dom,
                              "candidates_per_particle": cand, "bytes_per_particle": 16.0 * cand,
                              "GB/s": 16.0 * cand * n / t_dom / 1e9 if dom == "density" else None,
                              "note": "the pair kernels load a candidate once for two targets, so about half of this crosses the L1 data pipe"}
    cb = None
    if not args.no_cpu_baseline:
        if grid is not None:
            # the settled state, so the CPU baseline also has particles in contact with the terrain
            th = grid.heights()
            cb = cpu_port_baseline(sim.download("pos"), L, gy, terrain=(th, tinfo["terrain_origin"], tinfo["terrain_cell"], tinfo["erosion"]),
                                   vel=sim.download("vel"))
        else:
            cb = cpu_port_baseline(pos, L, gy)
    refg = None
    if not args.gravity_unscaled and not args.no_reference_gravity:
        del sim2
        refg = reference_gravity_run(args, pkg, local, pos, L, n_axis, terrain)
    ns_total = nbr_end
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "particles": n, "h": 0.0457, "spacing": SPACING,
                       "dt": 0.01, "box_half_extent": L, "gravity_y": gy,
                       "gravity_note": "g scaled by 10/n_axis: dynamic similarity with the reference default scene (bench.scene_gravity); the run with the "
                                       "reference's g = -9.82 is under reference_gravity" if not args.gravity_unscaled else "unscaled g (the reference default)",
                       "mean_neighbours_after_run": (ns_total / n) if ns_total else None,
                       "l2": ("flushed between timed steps (%d MiB memset)" % args.l2_flush_mib) if args.l2_flush_mib else
                             "not flushed: the working set (%.0f MB of particle arrays + neighbour lists) exceeds the 126 MB L2" % (n * 400 / 1e6),
                       "timing": "one CUDA-event pair around all %d steps on the launching stream (same method at every N)" % args.steps,
                       "density_variant": args.density_variant, "force_variant": args.force_variant,
                       "neighbour_list_rows": rows_end, "neighbour_list_smem_entries": smem_end,
                       "list_overflow_pairs_last_step": ovf_end, **tinfo},
            "parity_sampled": gate_ok, "parity_gate": gate,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cb,
            "reference_gravity": refg}
    emit(line)


def quiet_stdout():
    """STDOUT must carry ONE JSON line, but libraries write banners there (NCCL prints its version on fd 1):
    point fd 1 at stderr for the whole run and keep the real stdout for emit().  The saved descriptor lives
    in the environment because slabs.py imports this file as a second module (`bench` vs `__main__`)."""
    if "SPHE_BENCH_JSON_FD" not in os.environ:
        sys.stdout.flush()
        os.environ["SPHE_BENCH_JSON_FD"] = str(os.dup(1))
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    fd = os.environ.get("SPHE_BENCH_JSON_FD")
    if fd is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(int(fd), data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50, help="timed steps")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS),
                    help="c3 (default) = BASELINE configs[2], the largest single-GPU configuration; under torchrun the same default is 4.096M particles per GPU = configs[3] at N = 8")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--density-variant", type=int, default=6, help="6 = pair index lists, software-prefetched candidate stream (default); 3 = plain lists; 20 = TMA-staged candidates in shared memory + bit-mask lists; 0 = thread per particle")
    ap.add_argument("--force-variant", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo/migration transport: peer = pack kernel stores into the neighbours' mailboxes over NVLink (default); nccl = NCCL P2P group")
    ap.add_argument("--zone-sums", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU slab-local terrain: transport of the boundary-zone sums (peer = through the slab mailboxes, one launch per sum; nccl = P2P groups)")
    ap.add_argument("--terrain-share", default="window", choices=["window", "allreduce"],
                    help="multi-GPU terrain: window = slab-local rows + boundary-zone sums with the x-neighbours (default); allreduce = full replicas, 2 all-reduces per step")
    ap.add_argument("--rebalance-every", type=int, default=0, help="multi-GPU: re-cut the slabs by particle count every K steps; a slab-local terrain moves its row windows with the cuts (0 = never; the bench scenes are balanced by construction)")
    ap.add_argument("--layout", default="auto", choices=["auto", "tiled", "contiguous"], help="multi-GPU scene layout (slabs.channel_block)")
    ap.add_argument("--slab-lag", type=int, default=2, help="multi-GPU: steps the host may run ahead (0 = one host sync per step)")
    ap.add_argument("--settle", type=int, default=150, help="untimed steps before warm-up when the workload has a terrain")
    ap.add_argument("--gravity-unscaled", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true", help="skip the full-size one-step comparison with the oracle after the timed region")
    ap.add_argument("--no-reference-gravity", action="store_true", help="skip the second, labelled run with the reference's g = -9.82")
    ap.add_argument("--l2-flush-mib", type=int, default=0, help="flush L2 with a memset of this size between the steps of the instrumented pass (small workloads only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
