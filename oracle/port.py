"""ctypes wrapper over oracle/_build/libsphoracle.so (oracle/sph_oracle.c).

TEST INFRASTRUCTURE ONLY: the plain-C restatement of the reference hot path.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libsphoracle.so")


class Params(C.Structure):
    _fields_ = [("mass", C.c_float), ("visc", C.c_float), ("surf_tens", C.c_float), ("p0", C.c_float),
                ("k", C.c_float), ("h", C.c_float), ("len", C.c_float), ("dt", C.c_float),
                ("g", C.c_float * 3)]


class Grid(C.Structure):
    _fields_ = [("gmin", C.c_float * 3), ("cell", C.c_float), ("dim", C.c_int * 3)]

    @property
    def ncells(self):
        return int(self.dim[0]) * int(self.dim[1]) * int(self.dim[2])


class _State(C.Structure):
    _fields_ = [("n", C.c_int)] + [(k, C.c_void_p) for k in
                                   ("pos", "vel", "acc", "density", "pressure", "fpress", "fvisc", "fgrav",
                                    "fsurf", "normal", "neighb")]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("sph_oracle.c", "sph_oracle.h", "terrain_oracle.c", "terrain_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.so_lattice.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.so_step_allpairs.argtypes = [C.c_void_p, C.c_void_p]
        L.so_step_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.so_grid_for_box.argtypes = [C.c_void_p] * 4
        L.so_forces_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.so_integrate.argtypes = [C.c_void_p] * 4
        L.so_box.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.so_bin.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.so_neighbours.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.so_neighbours.restype = C.c_long
        L.so_default_params.argtypes = [C.c_void_p]
        L.so_terrain_collision.argtypes = [C.c_void_p] * 6
        L.so_terrain_surface.argtypes = [C.c_void_p, C.c_void_p]
        L.so_terrain_indices.argtypes = [C.c_void_p, C.c_void_p]
        L.so_terrain_indices.restype = C.c_long
        L.so_terrain_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_float, C.c_float, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(**kw):
    P = Params()
    lib().so_default_params(C.byref(P))
    for k, v in kw.items():
        if k == "g":
            P.g[:] = list(v)
        else:
            setattr(P, k, v)
    return P


def lattice(n, origin=(0.0, 0.0, 0.0)):
    o = np.asarray(origin, np.float32)
    cnt = lib().so_lattice(n, _p(o), None)
    pos = np.zeros((cnt, 3), np.float32)
    lib().so_lattice(n, _p(o), _p(pos))
    return pos


class State:
    """SoA particle state + diagnostics, in particle-id order."""
    V3 = ("pos", "vel", "acc", "fpress", "fvisc", "fgrav", "fsurf", "normal")
    S1 = ("density", "pressure")

    def __init__(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        self.n = n
        self.pos = pos.copy()
        self.vel = np.zeros((n, 3), np.float32) if vel is None else np.ascontiguousarray(vel, np.float32).copy()
        for k in self.V3[2:]:
            setattr(self, k, np.zeros((n, 3), np.float32))
        for k in self.S1:
            setattr(self, k, np.zeros(n, np.float32))
        self.neighb = np.zeros(n, np.int32)

    def _c(self):
        s = _State()
        s.n = self.n
        for k in self.V3 + self.S1 + ("neighb",):
            setattr(s, k, _p(getattr(self, k)))
        return s


def step_allpairs(P, S, steps=1):
    cs = S._c()
    for _ in range(steps):
        lib().so_step_allpairs(C.byref(P), C.byref(cs))


def grid_for_box(P, lo, hi):
    G = Grid()
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    lib().so_grid_for_box(C.byref(P), _p(lo), _p(hi), C.byref(G))
    return G


def step_grid(P, G, S, steps=1):
    cs = S._c()
    for _ in range(steps):
        lib().so_step_grid(C.byref(P), C.byref(G), C.byref(cs))


def step_grid_terrain(P, G, S, T, E, sediment, cR=0.5):
    """One step with a terrain attached: passes 1-3, integration, terrain stage (contact response +
    erosion, oracle/terrain_oracle.c), then the box collision -- the order of fluid_system.h:335-347
    with the commented call live.  sediment: int32 fixed point, updated in place.  Returns hit flags."""
    cs = S._c()
    lib().so_forces_grid(C.byref(P), C.byref(G), C.byref(cs))
    pn = np.zeros_like(S.pos); vn = np.zeros_like(S.vel)
    lib().so_integrate(C.byref(P), C.byref(cs), _p(pn), _p(vn))
    hit = T.stage(E, S.pos, pn, vn, sediment, P.dt, cR)
    lib().so_box(C.byref(P), S.n, _p(pn), _p(vn))
    S.pos[:] = pn; S.vel[:] = vn
    return hit


def bin_particles(G, pos):
    pos = np.ascontiguousarray(pos, np.float32)
    n = pos.shape[0]
    cell_of = np.zeros(n, np.int32); order = np.zeros(n, np.int32)
    cell_start = np.zeros(G.ncells + 1, np.int32)
    lib().so_bin(C.byref(G), n, _p(pos), _p(cell_of), _p(order), _p(cell_start))
    return cell_of, order, cell_start


def neighbours(P, G, pos, order, cell_start):
    pos = np.ascontiguousarray(pos, np.float32)
    n = pos.shape[0]
    ns = np.zeros(n + 1, np.int64)
    total = lib().so_neighbours(C.byref(P), C.byref(G), n, _p(pos), _p(order), _p(cell_start), _p(ns), None)
    nb = np.zeros(max(total, 1), np.int32)
    lib().so_neighbours(C.byref(P), C.byref(G), n, _p(pos), _p(order), _p(cell_start), _p(ns), _p(nb))
    return ns, nb[:total]


def omp_threads():
    return lib().so_omp_threads()


# ----------------------------------------------------------------------------- terrain (terrain_oracle.c)
FX = 4096  # fixed-point scale of heights and sediment


class _Terrain(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("dimx", C.c_int), ("dimy", C.c_int), ("dimz", C.c_int),
                ("h", C.c_void_p), ("hfx", C.c_void_p)]


class Erosion(C.Structure):
    _fields_ = [("enabled", C.c_int), ("origin", C.c_float * 3), ("scale", C.c_float), ("Kc", C.c_float),
                ("Ke", C.c_float), ("Kd", C.c_float), ("hmin_fx", C.c_int), ("max_pickup_fx", C.c_int)]


def erosion_params(enabled=True, origin=(0.0, 0.0, 0.0), scale=1.0, Kc=0.05, Ke=0.3, Kd=0.3, hmin=0.0, max_pickup=0.25):
    E = Erosion()
    E.enabled = int(enabled); E.origin[:] = list(origin); E.scale = scale
    E.Kc, E.Ke, E.Kd = Kc, Ke, Kd
    E.hmin_fx = int(round(hmin * FX)); E.max_pickup_fx = int(round(max_pickup * FX))
    return E


class Terrain:
    """Heightfield (rows x cols, H(x, z) = h[x, z]) + Grid dims, as Erosion/grid.h holds them."""

    def __init__(self, heights, dims):
        h = np.ascontiguousarray(heights, np.float32)
        self.hfx = np.ascontiguousarray(np.rint(h.astype(np.float64) * FX), np.int32)
        self.h = (self.hfx.astype(np.float32) * np.float32(1.0 / FX)).astype(np.float32)
        self.dims = tuple(int(d) for d in dims)

    def _c(self):
        t = _Terrain()
        t.rows, t.cols = self.h.shape
        t.dimx, t.dimy, t.dimz = self.dims
        t.h = _p(self.h); t.hfx = _p(self.hfx)
        return t

    def collision(self, pc, pn, vn):
        pc = np.ascontiguousarray(pc, np.float32); pn = np.ascontiguousarray(pn, np.float32)
        vn = np.ascontiguousarray(vn, np.float32)
        n = pc.shape[0]
        hit = np.zeros(n, np.int32); cp = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
        t = self._c()
        f = lib().so_terrain_collision
        for i in range(n):
            hit[i] = f(C.byref(t), _p(pc[i]), _p(pn[i]), _p(vn[i]), _p(cp[i]), _p(nrm[i]))
        return hit, cp, nrm

    def surface(self):
        out = np.zeros(self.dims[0] * self.dims[2] * 6, np.float32)
        t = self._c()
        lib().so_terrain_surface(C.byref(t), _p(out))
        return out

    def indices(self):
        t = self._c()
        k = lib().so_terrain_indices(C.byref(t), None)
        out = np.zeros(k, np.uint32)
        lib().so_terrain_indices(C.byref(t), _p(out))
        return out

    def stage(self, E, pos_curr, pos_next, vel_next, sediment, dt, cR=0.5):
        """Terrain stage of the step (contact response + erosion); arrays are updated in place."""
        n = pos_curr.shape[0]
        hit = np.zeros(n, np.int32)
        t = self._c()
        lib().so_terrain_stage(C.byref(t), C.byref(E), n, _p(pos_curr), _p(pos_next), _p(vel_next), _p(sediment),
                               C.c_float(dt), C.c_float(cR), _p(hit))
        return hit
