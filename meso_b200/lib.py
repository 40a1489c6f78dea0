"""ctypes binding of include/meso_b200.h (the C-ABI drop-in boundary).

The product path: there is no CPU fallback -- if libmeso_b200.so is missing or no
sm_100 GPU is visible, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmeso_b200.so")

# every symbol include/meso_b200.h declares: (restype, argtypes)
_vp, _i, _d, _i64, _u64 = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_uint64
_pi, _pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
SIGNATURES = {
    "meso_device_count": (_i, []),
    "meso_create": (_i, [C.POINTER(_vp), _i]),
    "meso_create_gang": (_i, [C.POINTER(_vp), _i, _pi]),
    "meso_gang_size": (_i, [_vp]),
    "meso_gang_layout": (_i, [_i, _pd, _pd, _pi, _pi, _i, _vp, _vp]),
    "meso_destroy": (None, [_vp]),
    "meso_last_error": (C.c_char_p, [_vp]),
    "meso_sync": (_i, [_vp]),
    "meso_stream": (_vp, [_vp]),
    "meso_profiler": (_i, [_vp, _i]),
    "meso_memory_usage": (_i, [_vp, C.POINTER(_u64)]),
    "meso_set_box": (_i, [_vp, _pd, _pd, _pi]),
    "meso_comm_unique_id": (_i, [_vp]),
    "meso_set_decomposition": (_i, [_vp, _i, _pi, _vp]),
    "meso_comm_blob_size": (_i, []),
    "meso_comm_export": (_i, [_vp, _vp]),
    "meso_comm_import": (_i, [_vp, _vp, _i]),
    "meso_set_neighbor": (_i, [_vp, _d, _i]),
    "meso_set_types": (_i, [_vp, _i, _pd]),
    "meso_pair_dpd_settings": (_i, [_vp, _i, _d, _i]),
    "meso_pair_dpd_coeff": (_i, [_vp, _pd]),
    "meso_set_timestep_size": (_i, [_vp, _d]),
    "meso_set_force_units": (_i, [_vp, _d]),
    "meso_set_reduce_scope": (_i, [_vp, _i]),
    "meso_host_register": (_i, [_vp, _vp, _u64]),
    "meso_host_unregister": (_i, [_vp, _vp]),
    "meso_set_ntimestep": (_i, [_vp, _i64]),
    "meso_get_ntimestep": (_i64, [_vp]),
    "meso_atoms_upload": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "meso_atoms_download": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "meso_counts": (_i, [_vp, _pi, _pi, _pi, _pi]),
    "meso_natoms_global": (_i64, [_vp]),
    "meso_initial_integrate": (_i, [_vp, _i]),
    "meso_neighbor_decide": (_i, [_vp]),
    "meso_rebuild": (_i, [_vp]),
    "meso_forward_comm": (_i, [_vp]),
    "meso_force_clear": (_i, [_vp, _i, _i]),
    "meso_pair_compute": (_i, [_vp, _i, _i, _i]),
    "meso_final_integrate": (_i, [_vp, _i]),
    "meso_compute_ke": (_i, [_vp, _i, _pd, _pd]),
    "meso_compute_virial": (_i, [_vp, _pd, _pd]),
    "meso_bond_harmonic_coeff": (_i, [_vp, _i, _pd, _pd]),
    "meso_set_special_bonds": (_i, [_vp, _d]),
    "meso_bonds_upload": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i]),
    "meso_bond_compute": (_i, [_vp, _i, _i]),
    "meso_compute_bond_energy": (_i, [_vp, _pd]),
    "meso_fix_wall": (_i, [_vp, _i, _i, _d, _d]),
    "meso_fix_solid_bound": (_i, [_vp, _i, _i, _i]),
    "meso_fix_addforce": (_i, [_vp, _i, _d, _d, _d]),
    "meso_fix_pois": (_i, [_vp, _i, _i, _i, _d, _d]),
    "meso_fix_rdf": (_i, [_vp, _i, _i, _i, _i]),
    "meso_fix_rdf_read": (_i, [_vp, _i, _i, _pd, _pd, _pd, _pd]),
    "meso_fix_clear": (_i, [_vp]),
    "meso_fix_post_force": (_i, [_vp, _i]),
    "meso_fix_bounce": (_i, [_vp, _i]),
    "meso_setup": (_i, [_vp, _i, _i]),
    "meso_run": (_i, [_vp, _i, _i]),
    "meso_export_bins": (_i, [_vp, _pi, _pd, _pd, _pi]),
    "meso_export_reorder": (_i, [_vp, _i, _vp, _vp]),
    "meso_export_packed": (_i, [_vp, _i, _vp, _vp]),
    "meso_export_ghosts": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "meso_export_cells": (_i, [_vp, _i, _vp, _i, _vp]),
    "meso_export_stencil": (_i, [_vp, _i, _vp]),
    "meso_export_pair_count": (_i, [_vp, _i, _vp]),
    "meso_export_pair_table": (_i, [_vp, _i64, _vp]),
    "meso_export_pair_rows": (_i, [_vp, _i64, _vp, _vp]),
    "meso_export_virial": (_i, [_vp, _i, _vp, _vp]),
    "meso_eval_gaussian": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "meso_eval_math": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "meso_eval_log2u": (_i, [_vp, _i, _vp, _vp]),
    "meso_timers_enable": (_i, [_vp, _i]),
    "meso_timers_read": (_i, [_vp, _pd, C.POINTER(_i64), _i]),
    "meso_launch_count": (_i, [_vp, C.POINTER(_i64), _i]),
}

MESO_BULK, MESO_BORDER, MESO_LOCAL, MESO_GHOST, MESO_ALL = 1, 2, 3, 4, 7
MESO_SP, MESO_DP = 0, 1
TIMER_NAMES = ("integrate", "forward", "pair", "rebuild", "neigh")

_lib = None


def load():
    """dlopen the in-tree library and type every entry point.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)          # AttributeError here == header and library disagree
        f.restype, f.argtypes = res, args
    _lib = L
    return L
