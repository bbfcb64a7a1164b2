"""Python mirror of the reference's `FluidSystemSPH` public surface (Erosion/fluid_system.h:66-289)
over the C ABI -- same method names and argument meaning, used by tests/, bench.py and the multi-GPU
driver.  The C++ drop-in shim with the identical surface is host/fluid_system.h."""
import ctypes as C

import numpy as np

from . import capi


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class FluidSystemSPH:
    def __init__(self, device=None):
        self._L = capi.lib()
        h = C.c_void_p()
        capi.check(self._L.sphe_create(C.byref(h)))  # no CUDA work here (fluid_system.h:69-72)
        self._h = h
        if device is not None:
            capi.check(self._L.sphe_set_device(h, int(device)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.sphe_destroy(self._h)
            self._h = None

    # ---- reference surface
    def Initialize(self, nParts): capi.check(self._L.sphe_initialize(self._h, int(nParts)))
    def AddParticles(self, n): capi.check(self._L.sphe_add_particles(self._h, int(n)))
    def Reset(self): capi.check(self._L.sphe_reset(self._h))

    def Run(self, grid=None):
        capi.check(self._L.sphe_step(self._h, getattr(grid, "_t", None)))

    def SetOrigin(self, o):
        a = np.asarray(o, np.float32)
        capi.check(self._L.sphe_set_origin(self._h, _p(a)))

    def GetOrigin(self):
        a = np.zeros(3, np.float32)
        capi.check(self._L.sphe_get_origin(self._h, _p(a)))
        return a

    def SetDeltaTime(self, dt): capi.check(self._L.sphe_set_dt(self._h, float(dt)))
    def GetDeltaTime(self): return self._L.sphe_get_dt(self._h)

    @property
    def params(self):
        """The host-resident parameter block behind GetMass/GetVisc/GetSurfTen/Getp0/GetGrav."""
        return self._L.sphe_params_ptr(self._h).contents

    def GetParticle(self, id):
        p = capi.Particle()
        capi.check(self._L.sphe_get_particle(self._h, int(id), C.byref(p)))
        return p

    def PrintCoords(self):
        pos = self.download("pos")
        for i in range(min(self._L.sphe_num(self._h), pos.shape[0])):
            print("[%d] %g %g %g" % (i, pos[i, 0], pos[i, 1], pos[i, 2]))

    # ---- extras of the C ABI
    def count(self): return self._L.sphe_count(self._h)
    def sync(self): capi.check(self._L.sphe_sync(self._h))
    def set_diagnostics(self, on=True): capi.check(self._L.sphe_set_diagnostics(self._h, int(on)))
    def set_variant(self, density=6, force=3): capi.check(self._L.sphe_set_variant(self._h, density, force))

    def set_grid_bounds(self, lo, hi):
        lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
        capi.check(self._L.sphe_set_grid_bounds(self._h, _p(lo), _p(hi)))

    def grid_info(self):
        g = capi.GridInfo()
        capi.check(self._L.sphe_grid_info_get(self._h, C.byref(g)))
        return g

    def upload_state(self, pos, vel):
        pos = np.ascontiguousarray(pos, np.float32); vel = np.ascontiguousarray(vel, np.float32)
        assert pos.shape == vel.shape and pos.ndim == 2 and pos.shape[1] == 3
        capi.check(self._L.sphe_upload_state(self._h, pos.shape[0], _p(pos), _p(vel)))

    def download(self, name):
        fid, w, dt = capi.FIELDS[name]
        n = self.count()
        out = np.zeros((n, w) if w > 1 else (n,), np.dtype(dt))
        capi.check(self._L.sphe_download(self._h, fid, _p(out)))
        return out

    def step_host_ptr(self, n, pos_in, vel_in, pos_out, vel_out, rho_out=None, grid=None):
        """End-to-end host-buffer step; arguments are raw host addresses (ints), e.g. pinned tensors."""
        capi.check(self._L.sphe_step_host(self._h, getattr(grid, "_t", None), int(n), pos_in, vel_in, pos_out, vel_out, rho_out))

    def step_host(self, pos, vel, grid=None):
        pos = np.ascontiguousarray(pos, np.float32); vel = np.ascontiguousarray(vel, np.float32)
        n = pos.shape[0]
        po = np.empty_like(pos); vo = np.empty_like(vel); rho = np.empty(n, np.float32)
        capi.check(self._L.sphe_step_host(self._h, getattr(grid, "_t", None), n, _p(pos), _p(vel), _p(po), _p(vo), _p(rho)))
        return po, vo, rho

    def write_positions_device(self, device_ptr, capacity_floats):
        """Packed xyz positions (id order) into a caller-owned device buffer, e.g. a mapped GL vertex buffer."""
        capi.check(self._L.sphe_write_positions_device(self._h, C.c_void_p(int(device_ptr)), int(capacity_floats)))

    def set_l2_flush(self, nbytes): capi.check(self._L.sphe_set_l2_flush(self._h, int(nbytes)))

    def timed_steps(self, steps, grid=None, per_kernel=True):
        ms = C.c_float(0)
        mk = (C.c_float * len(capi.K_NAMES))()
        nl = C.c_int(0)
        capi.check(self._L.sphe_timed_steps(self._h, getattr(grid, "_t", None), int(steps), C.byref(ms),
                                            mk if per_kernel else None, C.byref(nl)))
        return ms.value, dict(zip(capi.K_NAMES, [float(x) for x in mk])), nl.value

    def terrain_survivors(self):
        out = (C.c_int * 3)()
        capi.check(self._L.sphe_terrain_survivors(self._h, out))
        return [int(x) for x in out]

    def sediment_total_fx(self):
        v = C.c_longlong(0)
        capi.check(self._L.sphe_sediment_total_fx(self._h, C.byref(v)))
        return v.value

    def set_sediment_fx(self, sed_by_id):
        a = np.ascontiguousarray(sed_by_id, np.int32)
        assert a.shape[0] == self.count()
        capi.check(self._L.sphe_set_sediment_fx(self._h, _p(a)))

    def set_box(self, half):
        """Per-axis box half-extents (None restores the reference's cube `len`)."""
        a = None if half is None else np.asarray(half, np.float32)
        capi.check(self._L.sphe_set_box(self._h, _p(a)))

    def set_stream(self, cuda_stream):
        """Run on a caller-owned cudaStream_t (int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        capi.check(self._L.sphe_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def kernel_timing(self, on=True): capi.check(self._L.sphe_kernel_timing(self._h, int(on)))

    def kernel_times(self):
        mk = (C.c_float * len(capi.K_NAMES))()
        nl = C.c_int(0)
        capi.check(self._L.sphe_kernel_times(self._h, mk, C.byref(nl)))
        return dict(zip(capi.K_NAMES, [float(x) for x in mk])), nl.value

    # ---- multi-GPU x-slabs (include/sphe.h "multi-GPU x-slabs")
    def nlist_capacity(self):
        return int(self._L.sphe_nlist_capacity(self._h))

    def save_state(self, path, grid=None):
        """Checkpoint: particles (+ the terrain when given) to one file (sphe_save_state)."""
        capi.check(self._L.sphe_save_state(self._h, getattr(grid, "_t", None), str(path).encode()))

    def load_state(self, path, grid=None):
        capi.check(self._L.sphe_load_state(self._h, getattr(grid, "_t", None), str(path).encode()))

    def nlist_smem_entries(self):
        return int(self._L.sphe_nlist_smem_entries(self._h))

    def nlist_overflowed(self):
        return int(self._L.sphe_nlist_overflowed(self._h))

    def set_nlist_capacity(self, entries):
        capi.check(self._L.sphe_set_nlist_capacity(self._h, int(entries)))

    def slab_configure(self, x0, x1, has_left, has_right):
        capi.check(self._L.sphe_slab_configure(self._h, int(x0), int(x1), int(bool(has_left)), int(bool(has_right))))

    def slab_ring(self, wrap_left, wrap_right, far_x0=0):
        capi.check(self._L.sphe_slab_ring(self._h, int(bool(wrap_left)), int(bool(wrap_right)), int(far_x0)))

    def slab_info(self):
        v = [C.c_int(0) for _ in range(4)]
        capi.check(self._L.sphe_slab_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("gnx", "xoff", "n_total", "n_owned"), [x.value for x in v]))

    def slab_transit(self):
        out = (C.c_int * 3)()
        capi.check(self._L.sphe_slab_transit(self._h, out))
        return dict(zip(("to_left", "to_right", "forwarded"), list(out)))

    def slab_upload(self, pos, vel, ids):
        pos = np.ascontiguousarray(pos, np.float32); vel = np.ascontiguousarray(vel, np.float32)
        ids = np.ascontiguousarray(ids, np.int32)
        capi.check(self._L.sphe_slab_upload(self._h, pos.shape[0], _p(pos), _p(vel), _p(ids)))

    def slab_upload_ptr(self, n, pos, vel, ids):
        capi.check(self._L.sphe_slab_upload(self._h, int(n), pos, vel, ids))

    def slab_pack(self, dev_left, dev_right, cap_records, reserve_incoming):
        capi.check(self._L.sphe_slab_pack(self._h, dev_left, dev_right, int(cap_records), int(reserve_incoming)))

    def slab_unpack(self, dev_left, max_left, dev_right, max_right):
        out = (C.c_int * 6)()
        capi.check(self._L.sphe_slab_unpack(self._h, dev_left, int(max_left), dev_right, int(max_right), out))
        return dict(zip(("n_total", "n_owned", "to_left", "to_right", "from_left", "from_right"), list(out)))

    def slab_unpack_async(self, dev_left, max_left, dev_right, max_right):
        t = C.c_longlong(0)
        capi.check(self._L.sphe_slab_unpack_async(self._h, dev_left, int(max_left), dev_right, int(max_right), C.byref(t)))
        return t.value

    def slab_result(self, ticket, wait=True):
        out = (C.c_int * 6)()
        rc = self._L.sphe_slab_result(self._h, int(ticket), int(bool(wait)), out)
        if rc == 1:
            return None
        capi.check(rc)
        return dict(zip(("n_total", "n_owned", "to_left", "to_right", "from_left", "from_right"), list(out)))

    # peer-memory exchange (include/sphe.h "Peer-memory exchange")
    def slab_zone_sum(self, grid, which, off_left, off_right, count):
        capi.check(self._L.sphe_slab_zone_sum(self._h, grid._t, int(which), int(off_left), int(off_right), int(count)))

    def slab_peer_setup_zones(self, cap_records, reserve_particles, zone_ints):
        capi.check(self._L.sphe_slab_peer_setup_zones(self._h, int(cap_records), int(reserve_particles), int(zone_ints)))

    def slab_peer_setup(self, cap_records, reserve_particles=0):
        capi.check(self._L.sphe_slab_peer_setup(self._h, int(cap_records), int(reserve_particles)))

    def slab_peer_handle(self):
        buf = C.create_string_buffer(64)
        capi.check(self._L.sphe_slab_peer_handle(self._h, buf))
        return buf.raw

    def slab_peer_connect(self, left_handle, right_handle):
        l = C.create_string_buffer(left_handle, 64) if left_handle is not None else None
        r = C.create_string_buffer(right_handle, 64) if right_handle is not None else None
        capi.check(self._L.sphe_slab_peer_connect(self._h, l, r))

    def slab_peer_connect_local(self, left, right):
        capi.check(self._L.sphe_slab_peer_connect_local(self._h, left._h if left is not None else None,
                                                        right._h if right is not None else None))

    def slab_peer_timeout(self, clock_cycles):
        capi.check(self._L.sphe_slab_peer_timeout(self._h, int(clock_cycles)))

    def slab_send(self):
        capi.check(self._L.sphe_slab_send(self._h))

    def slab_recv(self):
        t = C.c_longlong(0)
        capi.check(self._L.sphe_slab_recv(self._h, C.byref(t)))
        return t.value

    def step_phase(self, grid, phase):
        """One phase (0, 1, 2) of a step whose terrain erosion is shared by several slabs (sphe_step_phase)."""
        capi.check(self._L.sphe_step_phase(self._h, grid._t if grid is not None else None, int(phase)))

    def slab_column_histogram(self, gnx):
        out = np.zeros(int(gnx), np.int32)
        capi.check(self._L.sphe_slab_column_histogram(self._h, int(gnx), _p(out)))
        return out

    def slab_download(self, cap=None):
        cap = self.count() if cap is None else int(cap)
        ids = np.zeros(cap, np.int32); pos = np.zeros((cap, 3), np.float32); vel = np.zeros((cap, 3), np.float32)
        rho = np.zeros(cap, np.float32); sed = np.zeros(cap, np.float32)
        m = C.c_int(0)
        capi.check(self._L.sphe_slab_download(self._h, cap, _p(ids), _p(pos), _p(vel), _p(rho), _p(sed), C.byref(m)))
        m = m.value
        return ids[:m], pos[:m], vel[:m], rho[:m], sed[:m]

    def slab_download_ptr(self, cap, ids, pos, vel, rho):
        m = C.c_int(0)
        capi.check(self._L.sphe_slab_download(self._h, int(cap), ids, pos, vel, rho, None, C.byref(m)))
        return m.value

    # ---- neighbour-grid test hooks
    def debug_cells(self):
        out = np.zeros(self.count(), np.int32)
        capi.check(self._L.sphe_debug_cells(self._h, _p(out)))
        return out

    def debug_sorted_order(self):
        out = np.zeros(self.count(), np.int32)
        capi.check(self._L.sphe_debug_sorted_order(self._h, _p(out)))
        return out

    def debug_cell_start(self):
        g = self.grid_info()
        out = np.zeros(int(g.dim[0]) * int(g.dim[1]) * int(g.dim[2]) + 1, np.int32)
        capi.check(self._L.sphe_debug_cell_start(self._h, _p(out)))
        return out

    def debug_neighbours_total(self):
        tot = C.c_longlong(0)
        capi.check(self._L.sphe_debug_neighbours(self._h, None, None, 0, C.byref(tot)))
        return tot.value

    def debug_pair_lists(self, cap=512):
        """Decoded production neighbour masks of the last step: (counts[n], entries[n, cap]) by sorted slot."""
        n = self.count()
        counts = np.zeros(n, np.int32); entries = np.zeros((n, cap), np.int32)
        capi.check(self._L.sphe_debug_pair_lists(self._h, cap, _p(counts), _p(entries)))
        return counts, entries

    def debug_neighbours(self):
        n = self.count()
        ns = np.zeros(n + 1, np.int64)
        tot = C.c_longlong(0)
        capi.check(self._L.sphe_debug_neighbours(self._h, _p(ns), None, 0, C.byref(tot)))
        nb = np.zeros(max(tot.value, 1), np.int32)
        capi.check(self._L.sphe_debug_neighbours(self._h, _p(ns), _p(nb), tot.value, C.byref(tot)))
        return ns, nb[:tot.value]
