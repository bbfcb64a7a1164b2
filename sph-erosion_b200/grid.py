"""Python mirror of the reference's `Grid` public surface (Erosion/grid.h:72-841) over the C ABI
(include/sphe.h "terrain").  Same method names and argument meaning; everything runs on the GPU.
The C++ drop-in shim with the identical surface is host/grid.h."""
import ctypes as C

import numpy as np

from . import capi


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Grid:
    def __init__(self, dimX=512, dimY=512, dimZ=512):
        self._L = capi.lib()
        t = C.c_void_p()
        capi.check(self._L.sphe_terrain_create(C.byref(t), int(dimX), int(dimY), int(dimZ)))  # no CUDA work yet
        self._t = t

    def __del__(self):
        if getattr(self, "_t", None):
            self._L.sphe_terrain_destroy(self._t)
            self._t = None

    # ---- reference surface
    def LoadHeightfield(self, img):
        """512 x 512 unsigned char, H(x, z) = img[x, z] (grid.h:98-107).  Other sizes: load_heightfield."""
        img = np.ascontiguousarray(img, np.uint8)
        assert img.size == 512 * 512
        capi.check(self._L.sphe_terrain_load_heightfield(self._t, _p(img)))

    def GetHeightfieldAt(self, x, y):
        v = self._L.sphe_terrain_height_at(self._t, int(x), int(y))
        if v < 0:
            raise capi.SpheError(self._L.sphe_last_error().decode())
        return v

    def UpdateGrid(self, dimx, dimy, dimz):
        capi.check(self._L.sphe_terrain_update_grid(self._t, int(dimx), int(dimy), int(dimz)))

    def GetSurfacePartsSize(self): return self._L.sphe_terrain_surface_size(self._t)
    def GetIndicesSize(self): return self._L.sphe_terrain_indices_size(self._t)

    def GetSurfaceParts(self):
        out = np.zeros(self.GetSurfacePartsSize(), np.float32)
        capi.check(self._L.sphe_terrain_get_surface(self._t, _p(out)))
        return out

    def GetIndices(self):
        out = np.zeros(self.GetIndicesSize(), np.uint32)
        capi.check(self._L.sphe_terrain_get_indices(self._t, _p(out)))
        return out

    def GetDim(self):
        d = (C.c_int * 3)()
        capi.check(self._L.sphe_terrain_size(self._t, None, None, d))
        return tuple(d)

    def collision(self, posCurr, posNext, velNext):
        """Batched Grid::collision (grid.h:462-805): (n, 3) arrays in terrain coordinates ->
        (hit, contactP, norm)."""
        pc = np.ascontiguousarray(posCurr, np.float32).reshape(-1, 3)
        pn = np.ascontiguousarray(posNext, np.float32).reshape(-1, 3)
        vn = np.ascontiguousarray(velNext, np.float32).reshape(-1, 3)
        n = pc.shape[0]
        hit = np.zeros(n, np.int32); cp = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
        capi.check(self._L.sphe_terrain_collision(self._t, n, _p(pc), _p(pn), _p(vn), _p(hit), _p(cp), _p(nrm)))
        return hit, cp, nrm

    # ---- extras of the C ABI
    def load_heightfield(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        capi.check(self._L.sphe_terrain_load_heightfield_ex(self._t, _p(img), img.shape[0], img.shape[1]))

    def set_heights(self, h):
        h = np.ascontiguousarray(h, np.float32)
        capi.check(self._L.sphe_terrain_set_heights(self._t, _p(h), h.shape[0], h.shape[1]))

    def shape(self):
        r, c = C.c_int(0), C.c_int(0)
        capi.check(self._L.sphe_terrain_size(self._t, C.byref(r), C.byref(c), None))
        return r.value, c.value

    def heights(self):
        out = np.zeros(self.shape(), np.float32)
        capi.check(self._L.sphe_terrain_get_heights(self._t, _p(out)))
        return out

    def heights_fx(self):
        out = np.zeros(self.shape(), np.int32)
        capi.check(self._L.sphe_terrain_get_heights_fx(self._t, _p(out)))
        return out

    def set_transform(self, origin, scale):
        o = np.asarray(origin, np.float32)
        capi.check(self._L.sphe_terrain_set_transform(self._t, _p(o), float(scale)))

    @property
    def erosion(self):
        return self._L.sphe_terrain_erosion_ptr(self._t).contents

    def accumulators(self):
        """Device addresses of the per-vertex erosion accumulators (want, delta) and their length (int32 each):
        multi-GPU runs sum them over the ranks between the phases of a step (sphe_terrain_accumulators)."""
        w, d, n = C.c_void_p(0), C.c_void_p(0), C.c_longlong(0)
        capi.check(self._L.sphe_terrain_accumulators(self._t, C.byref(w), C.byref(d), C.byref(n)))
        return w.value, d.value, n.value

    def heights_device(self):
        """(device address, length) of the fixed-point heights (int32): re-cuts of a slab-local terrain sum the owners'
        rows into it (sphe_terrain_heights_device); call refresh() afterwards."""
        h, n = C.c_void_p(0), C.c_longlong(0)
        capi.check(self._L.sphe_terrain_heights_device(self._t, C.byref(h), C.byref(n)))
        return h.value, n.value

    def refresh(self):
        """Rebuild the cull map over the current window after the heights were written from outside (sphe_terrain_refresh)."""
        capi.check(self._L.sphe_terrain_refresh(self._t))

    def total_fx(self, rows=None):
        """Sum of the fixed-point heights (of the rows [rows[0], rows[1]) when given)."""
        v = C.c_longlong(0)
        if rows is None:
            capi.check(self._L.sphe_terrain_total_fx(self._t, C.byref(v)))
        else:
            capi.check(self._L.sphe_terrain_total_fx_rows(self._t, int(rows[0]), int(rows[1]), C.byref(v)))
        return v.value

    def set_window(self, row0, row1):
        """Slab-local terrain: keep only the rows [row0, row1) current on this replica (sphe_terrain_set_window)."""
        capi.check(self._L.sphe_terrain_set_window(self._t, int(row0), int(row1)))

    def window_violations(self):
        v = C.c_longlong(0)
        capi.check(self._L.sphe_terrain_window_violations(self._t, C.byref(v)))
        return v.value

    def contacts(self, reset=False):
        v = C.c_longlong(0)
        capi.check(self._L.sphe_terrain_contacts(self._t, C.byref(v), int(reset)))
        return v.value

    def stage(self, pos_curr, pos_next, vel_next, sediment, dt, cR=0.5):
        """Terrain stage on caller arrays (updated in place); returns the hit flags."""
        n = pos_curr.shape[0]
        hit = np.zeros(n, np.int32)
        capi.check(self._L.sphe_terrain_stage_host(self._t, n, _p(pos_curr), _p(pos_next), _p(vel_next), _p(sediment),
                                                   float(dt), float(cR), _p(hit)))
        return hit
