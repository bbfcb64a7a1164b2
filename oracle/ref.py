"""ctypes wrapper over oracle/_ref/libsphref{,_omp}.so -- the UNMODIFIED reference
(Erosion/fluid_system.h, Erosion/grid.h) compiled in place by oracle/Makefile.

TEST INFRASTRUCTURE ONLY.  The library is built in the authoring container
(where /root/reference exists) and travels to the GPU box as a prebuilt .so.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FIELDS = {"pos": (0, 3, np.float32), "vel": (1, 3, np.float32), "acc": (2, 3, np.float32),
          "density": (3, 1, np.float32), "pressure": (4, 1, np.float32),
          "fpress": (5, 3, np.float32), "fvisc": (6, 3, np.float32), "fgrav": (7, 3, np.float32),
          "fsurf": (8, 3, np.float32), "normal": (9, 3, np.float32),
          "id": (10, 1, np.int32), "neighb": (11, 1, np.int32)}


def lib_path(omp=False):
    return os.path.join(_HERE, "_ref", "libsphref_omp.so" if omp else "libsphref.so")


def available(omp=False):
    return os.path.exists(lib_path(omp))


_libs = {}


def _load(omp):
    if omp in _libs:
        return _libs[omp]
    L = C.CDLL(lib_path(omp))
    vp, f, i = C.c_void_p, C.c_float, C.c_int
    L.ref_create.restype = vp
    L.ref_grid_create.restype = vp
    L.ref_grid_create.argtypes = [i, i, i]
    L.ref_get_dt.restype = f
    L.ref_get_dt.argtypes = [vp]
    L.ref_grid_surface_size.restype = C.c_long
    L.ref_grid_indices_size.restype = C.c_long
    for name, args in {
        "ref_destroy": [vp], "ref_initialize": [vp, i], "ref_add_particles": [vp, i], "ref_reset": [vp],
        "ref_set_origin": [vp, f, f, f], "ref_set_dt": [vp, f], "ref_count": [vp],
        "ref_set_len": [vp, f], "ref_set_h": [vp, f], "ref_set_k": [vp, f],
        "ref_set_params": [vp, f, f, f, f, vp], "ref_get_params": [vp, vp],
        "ref_set_state": [vp, i, vp, vp], "ref_run": [vp, i], "ref_get_field": [vp, i, vp],
        "ref_grid_destroy": [vp], "ref_grid_load_heightfield": [vp, vp], "ref_grid_height_at": [vp, i, i],
        "ref_grid_update": [vp, i, i, i], "ref_grid_surface_size": [vp], "ref_grid_indices_size": [vp],
        "ref_grid_get_surface": [vp, vp], "ref_grid_get_indices": [vp, vp],
        "ref_grid_voxel_type": [vp, i, i, i],
        "ref_grid_collision": [vp, i, vp, vp, vp, vp, vp, vp],
    }.items():
        getattr(L, name).argtypes = args
    _libs[omp] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefSim:
    """The reference FluidSystemSPH driven headless."""

    def __init__(self, omp=False):
        self.L = _load(omp)
        self.h = C.c_void_p(self.L.ref_create())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    def initialize(self, n): self.L.ref_initialize(self.h, n)
    def add_particles(self, n): self.L.ref_add_particles(self.h, n)
    def reset(self): self.L.ref_reset(self.h)
    def set_origin(self, x, y, z): self.L.ref_set_origin(self.h, x, y, z)
    def set_dt(self, dt): self.L.ref_set_dt(self.h, dt)
    def get_dt(self): return self.L.ref_get_dt(self.h)
    def count(self): return self.L.ref_count(self.h)
    def set_len(self, v): self.L.ref_set_len(self.h, v)
    def set_h(self, v): self.L.ref_set_h(self.h, v)
    def set_k(self, v): self.L.ref_set_k(self.h, v)

    def set_params(self, mass, visc, surf, p0, g):
        g = np.asarray(g, np.float32)
        self.L.ref_set_params(self.h, mass, visc, surf, p0, _p(g))

    def params(self):
        out = np.zeros(10, np.float32)
        self.L.ref_get_params(self.h, _p(out))
        return dict(zip(["mass", "visc", "surf", "p0", "gx", "gy", "gz", "k", "h", "len"], out.tolist()))

    def set_state(self, pos, vel):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        self.L.ref_set_state(self.h, pos.shape[0], _p(pos), _p(vel))

    def run(self, steps=1): self.L.ref_run(self.h, steps)

    def field(self, name):
        fid, w, dt = FIELDS[name]
        n = self.count()
        out = np.zeros((n, w) if w > 1 else (n,), dt)
        self.L.ref_get_field(self.h, fid, _p(out))
        return out


class RefGrid:
    """The reference Grid (terrain) driven headless."""

    def __init__(self, dx, dy, dz, omp=False):
        self.L = _load(omp)
        self.g = C.c_void_p(self.L.ref_grid_create(dx, dy, dz))

    def __del__(self):
        if getattr(self, "g", None):
            self.L.ref_grid_destroy(self.g)
            self.g = None

    def load_heightfield(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        assert img.size == 512 * 512
        self.L.ref_grid_load_heightfield(self.g, _p(img))

    def height_at(self, x, y): return self.L.ref_grid_height_at(self.g, x, y)

    def update(self, dx, dy, dz): self.L.ref_grid_update(self.g, dx, dy, dz)

    def surface(self):
        out = np.zeros(self.L.ref_grid_surface_size(self.g), np.float32)
        self.L.ref_grid_get_surface(self.g, _p(out))
        return out

    def indices(self):
        out = np.zeros(self.L.ref_grid_indices_size(self.g), np.uint32)
        self.L.ref_grid_get_indices(self.g, _p(out))
        return out

    def voxel_type(self, x, y, z): return self.L.ref_grid_voxel_type(self.g, x, y, z)

    def collision(self, pc, pn, vn):
        pc = np.ascontiguousarray(pc, np.float32); pn = np.ascontiguousarray(pn, np.float32)
        vn = np.ascontiguousarray(vn, np.float32)
        n = pc.shape[0]
        hit = np.zeros(n, np.int32); cp = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
        self.L.ref_grid_collision(self.g, n, _p(pc), _p(pn), _p(vn), _p(hit), _p(cp), _p(nrm))
        return hit, cp, nrm


def fnv1a64(*arrays):
    """FNV-1a 64 over the little-endian bytes of per-particle interleaved fields."""
    cols = [np.ascontiguousarray(a).reshape(a.shape[0], -1).view(np.uint32) for a in arrays]
    data = np.concatenate(cols, axis=1).tobytes()
    h = 0xcbf29ce484222325
    for b in data:
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
