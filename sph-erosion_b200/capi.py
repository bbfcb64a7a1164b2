"""ctypes binding of include/sphe.h (libsphe_b200.so).  No compute happens here: every call goes
straight to the C ABI, and loading fails loudly if the CUDA library is missing."""
import ctypes as C
import os

from . import build as _build

K_NAMES = ["hash", "scan", "scatter", "reorder", "density", "force", "terrain"]
FIELDS = {"pos": (0, 3, "f4"), "vel": (1, 3, "f4"), "acc": (2, 3, "f4"), "density": (3, 1, "f4"),
          "pressure": (4, 1, "f4"), "fpress": (5, 3, "f4"), "fvisc": (6, 3, "f4"), "fgrav": (7, 3, "f4"),
          "fsurf": (8, 3, "f4"), "normal": (9, 3, "f4"), "id": (10, 1, "i4"), "neighb": (11, 1, "i4"),
          "sediment": (12, 1, "f4")}


class Params(C.Structure):
    _fields_ = [("mass", C.c_float), ("visc", C.c_float), ("surf_tens", C.c_float), ("p0", C.c_float),
                ("g", C.c_float * 3), ("dt", C.c_float), ("k", C.c_float), ("h", C.c_float),
                ("len", C.c_float), ("cR", C.c_float)]


class Particle(C.Structure):
    _fields_ = [("id", C.c_int), ("position", C.c_float * 3), ("velocity", C.c_float * 3),
                ("acceleration", C.c_float * 3), ("density", C.c_float), ("pressure", C.c_float),
                ("pressure_force", C.c_float * 3), ("viscosity_force", C.c_float * 3),
                ("gravity_force", C.c_float * 3), ("surface_force", C.c_float * 3),
                ("surface_normal", C.c_float * 3), ("neighb_id", C.c_int)]


class Erosion(C.Structure):
    _fields_ = [("enabled", C.c_int), ("Kc", C.c_float), ("Ke", C.c_float), ("Kd", C.c_float), ("hmin", C.c_float),
                ("max_pickup", C.c_float)]


class GridInfo(C.Structure):
    _fields_ = [("gmin", C.c_float * 3), ("cell", C.c_float), ("dim", C.c_int * 3)]


# every symbol include/sphe.h declares: name -> (restype, argtypes)
_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
SYMBOLS = {
    "sphe_last_error": (C.c_char_p, []),
    "sphe_abi_version": (_i, []),
    "sphe_create": (_i, [C.POINTER(_vp)]),
    "sphe_destroy": (None, [_vp]),
    "sphe_set_device": (_i, [_vp, _i]),
    "sphe_initialize": (_i, [_vp, _i]),
    "sphe_add_particles": (_i, [_vp, _i]),
    "sphe_reset": (_i, [_vp]),
    "sphe_set_origin": (_i, [_vp, _vp]),
    "sphe_get_origin": (_i, [_vp, _vp]),
    "sphe_set_dt": (_i, [_vp, _f]),
    "sphe_get_dt": (_f, [_vp]),
    "sphe_params_ptr": (C.POINTER(Params), [_vp]),
    "sphe_count": (_i, [_vp]),
    "sphe_num": (_i, [_vp]),
    "sphe_set_grid_bounds": (_i, [_vp, _vp, _vp]),
    "sphe_grid_info_get": (_i, [_vp, C.POINTER(GridInfo)]),
    "sphe_upload_state": (_i, [_vp, _i, _vp, _vp]),
    "sphe_step": (_i, [_vp, _vp]),
    "sphe_sync": (_i, [_vp]),
    "sphe_step_host": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sphe_set_l2_flush": (_i, [_vp, _ll]),
    "sphe_timed_steps": (_i, [_vp, _vp, _i, C.POINTER(_f), C.POINTER(_f), C.POINTER(_i)]),
    "sphe_set_diagnostics": (_i, [_vp, _i]),
    "sphe_get_particle": (_i, [_vp, _i, C.POINTER(Particle)]),
    "sphe_download": (_i, [_vp, _i, _vp]),
    "sphe_download_positions": (_i, [_vp, _vp]),
    "sphe_debug_cells": (_i, [_vp, _vp]),
    "sphe_debug_sorted_order": (_i, [_vp, _vp]),
    "sphe_debug_cell_start": (_i, [_vp, _vp]),
    "sphe_debug_neighbours": (_i, [_vp, _vp, _vp, _ll, C.POINTER(_ll)]),
    "sphe_debug_pair_lists": (_i, [_vp, _i, _vp, _vp]),
    "sphe_slab_column_histogram": (_i, [_vp, _i, _vp]),
    "sphe_terrain_survivors": (_i, [_vp, _vp]),
    "sphe_write_positions_device": (_i, [_vp, _vp, _ll]),
    "sphe_slab_peer_setup_zones": (_i, [_vp, _i, _i, _i]),
    "sphe_slab_zone_sum": (_i, [_vp, _vp, _i, _ll, _ll, _i]),
    "sphe_device_ptr": (_vp, [_vp, _i]),
    "sphe_set_stream": (_i, [_vp, _vp]),
    "sphe_set_variant": (_i, [_vp, _i, _i]),
    "sphe_set_box": (_i, [_vp, _vp]),
    "sphe_terrain_create": (_i, [C.POINTER(_vp), _i, _i, _i]),
    "sphe_terrain_destroy": (None, [_vp]),
    "sphe_terrain_load_heightfield": (_i, [_vp, _vp]),
    "sphe_terrain_load_heightfield_ex": (_i, [_vp, _vp, _i, _i]),
    "sphe_terrain_set_heights": (_i, [_vp, _vp, _i, _i]),
    "sphe_terrain_get_heights": (_i, [_vp, _vp]),
    "sphe_terrain_get_heights_fx": (_i, [_vp, _vp]),
    "sphe_terrain_size": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), _vp]),
    "sphe_terrain_height_at": (_i, [_vp, _i, _i]),
    "sphe_terrain_update_grid": (_i, [_vp, _i, _i, _i]),
    "sphe_terrain_surface_size": (_ll, [_vp]),
    "sphe_terrain_indices_size": (_ll, [_vp]),
    "sphe_terrain_get_surface": (_i, [_vp, _vp]),
    "sphe_terrain_get_indices": (_i, [_vp, _vp]),
    "sphe_terrain_collision": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sphe_terrain_set_transform": (_i, [_vp, _vp, _f]),
    "sphe_terrain_erosion_ptr": (C.POINTER(Erosion), [_vp]),
    "sphe_terrain_stage_host": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _f, _f, _vp]),
    "sphe_terrain_total_fx": (_i, [_vp, C.POINTER(_ll)]),
    "sphe_terrain_contacts": (_i, [_vp, C.POINTER(_ll), _i]),
    "sphe_terrain_total_fx_rows": (_i, [_vp, _i, _i, C.POINTER(_ll)]),
    "sphe_terrain_set_window": (_i, [_vp, _i, _i]),
    "sphe_terrain_window_violations": (_i, [_vp, C.POINTER(_ll)]),
    "sphe_sediment_total_fx": (_i, [_vp, C.POINTER(_ll)]),
    "sphe_set_sediment_fx": (_i, [_vp, _vp]),
    "sphe_kernel_timing": (_i, [_vp, _i]),
    "sphe_kernel_times": (_i, [_vp, C.POINTER(_f), C.POINTER(_i)]),
    "sphe_device_count": (_i, []),
    "sphe_nlist_capacity": (_i, [_vp]),
    "sphe_nlist_overflowed": (_i, [_vp]),
    "sphe_nlist_smem_entries": (_i, [_vp]),
    "sphe_set_nlist_capacity": (_i, [_vp, _i]),
    "sphe_slab_configure": (_i, [_vp, _i, _i, _i, _i]),
    "sphe_slab_ring": (_i, [_vp, _i, _i, _i]),
    "sphe_slab_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "sphe_slab_transit": (_i, [_vp, _vp]),
    "sphe_slab_upload": (_i, [_vp, _i, _vp, _vp, _vp]),
    "sphe_slab_pack": (_i, [_vp, _vp, _vp, _i, _i]),
    "sphe_slab_unpack": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "sphe_slab_unpack_async": (_i, [_vp, _vp, _i, _vp, _i, C.POINTER(_ll)]),
    "sphe_slab_result": (_i, [_vp, _ll, _i, _vp]),
    "sphe_slab_download": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i)]),
    "sphe_slab_peer_setup": (_i, [_vp, _i, _i]),
    "sphe_slab_peer_handle": (_i, [_vp, _vp]),
    "sphe_slab_peer_connect": (_i, [_vp, _vp, _vp]),
    "sphe_slab_peer_connect_local": (_i, [_vp, _vp, _vp]),
    "sphe_slab_peer_timeout": (_i, [_vp, _ll]),
    "sphe_slab_send": (_i, [_vp]),
    "sphe_slab_recv": (_i, [_vp, C.POINTER(_ll)]),
    "sphe_step_phase": (_i, [_vp, _vp, _i]),
    "sphe_save_state": (_i, [_vp, _vp, C.c_char_p]),
    "sphe_load_state": (_i, [_vp, _vp, C.c_char_p]),
    "sphe_terrain_accumulators": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_ll)]),
    "sphe_terrain_heights_device": (_i, [_vp, C.POINTER(_vp), C.POINTER(_ll)]),
    "sphe_terrain_refresh": (_i, [_vp]),
}

_lib = None


class SpheError(RuntimeError):
    pass


def lib():
    """Loads libsphe_b200.so; raises if it is missing (there is no fallback implementation)."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            raise SpheError("%s not found: run `python -c \"import __graft_entry__ as g; g.build()\"`" % path)
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise SpheError("sphe error %d: %s" % (rc, lib().sphe_last_error().decode()))
    return rc
