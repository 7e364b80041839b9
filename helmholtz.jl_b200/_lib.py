"""ctypes binding of libhelmholtz_b200.so (include/helmholtz_b200.h).

The library is the product; this module only declares its symbols.  There is no fallback: if the
shared object is missing or was built without CUDA the import raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libhelmholtz_b200.so")

HH_OK = 0
HH_NOT_CONVERGED = 1
HH_ERR_ARG, HH_ERR_CUDA, HH_ERR_STATE, HH_ERR_NAN, HH_ERR_UNSUPPORTED, HH_ERR_ALLOC = -1, -2, -3, -4, -5, -6
HH_C64, HH_C32, HH_C64_MIXED = 0, 1, 2
HH_RELAX_JAC, HH_RELAX_JAC_GMRES = 0, 1
HH_CYCLE_V, HH_CYCLE_W, HH_CYCLE_K = 0, 1, 2
HH_COARSE_LU, HH_COARSE_GMRES = 0, 1
HH_KRYLOV_GMRES, HH_KRYLOV_BICGSTAB = 0, 1
HH_MAX_LEVELS = 12


class hh_mg_options(C.Structure):
    _fields_ = [
        ("levels", C.c_int32),
        ("relax_type", C.c_int32),
        ("cycle_type", C.c_int32),
        ("coarse_type", C.c_int32),
        ("coarse_iters", C.c_int32),
        ("do_transpose", C.c_int32),
        ("relax_pre", C.c_int32 * HH_MAX_LEVELS),
        ("relax_post", C.c_int32 * HH_MAX_LEVELS),
        ("relax_param", C.c_double),
        ("shift", C.c_double * HH_MAX_LEVELS),
    ]


class hh_solve_options(C.Structure):
    _fields_ = [
        ("krylov", C.c_int32),
        ("inner", C.c_int32),
        ("max_iter", C.c_int32),
        ("do_transpose", C.c_int32),
        ("rel_tol", C.c_double),
    ]


class HelmholtzB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libhelmholtz_b200 error {code}: {msg}")
        self.code = code


_p = C.c_void_p
_i64p = C.POINTER(C.c_int64)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/helmholtz_b200.h declares
SIGNATURES = {
    "hh_version": (C.c_int, []),
    "hh_last_error": (C.c_char_p, [_p]),
    "hh_device_count": (C.c_int, [_ip]),
    "hh_get_abl": (C.c_int, [C.c_int, _i64p, C.c_int, _i64p, C.c_double, _dp]),
    "hh_get_maximal_frequency": (C.c_int, [_dp, C.c_int64, C.c_int, _dp, _dp]),
    "hh_point_source_index": (C.c_int64, [C.c_int, _i64p, _i64p]),
    "hh_ho_stencil": (C.c_int, [C.c_int, _i64p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp]),
    "hh_stencil_adjoint": (C.c_int, [C.c_int, _i64p, _dp, _dp]),
    "hh_assemble_csc": (C.c_int, [C.c_int, _i64p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                  _dp, _i64p, _i64p, _dp]),
    "hh_create": (C.c_int, [C.c_int, _i64p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                            C.c_int, C.POINTER(_p)]),
    "hh_create_multi": (C.c_int, [C.c_int, _i64p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_int, _ip, C.c_int, C.POINTER(_p)]),
    "hh_destroy": (C.c_int, [_p]),
    "hh_create_slab_local": (C.c_int, [C.c_int, _i64p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                       C.c_int, _ip, C.c_int, C.c_int, C.POINTER(_p)]),
    "hh_nccl_unique_id": (C.c_int, [_p]),
    "hh_create_slab_nccl": (C.c_int, [C.c_int, _i64p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_int64, C.c_int64, C.POINTER(_p)]),
    "hh_slab_info": (C.c_int, [_p, _ip, _ip, _ip, _i64p, _i64p]),
    "hh_slab_partition": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, _i64p]),
    "hh_slab_level_stencil": (C.c_int, [_p, C.c_int, C.c_int, _i64p, _p]),
    "hh_set_stream": (C.c_int, [_p, _p]),
    "hh_update_model": (C.c_int, [_p, _dp, _dp, C.c_double, C.c_double]),
    "hh_set_frequency_abl": (C.c_int, [_p, C.c_double, C.c_double, C.c_double, _i64p, C.c_double]),
    "hh_get_gamma": (C.c_int, [_p, _dp]),
    "hh_get_maximal_frequency_device": (C.c_int, [_p, _dp]),
    "hh_set_operator_ho": (C.c_int, [_p, C.c_int, _dp, _dp, _dp]),
    "hh_setup": (C.c_int, [_p, C.POINTER(hh_mg_options)]),
    "hh_clear": (C.c_int, [_p]),
    "hh_hierarchy_exists": (C.c_int, [_p]),
    "hh_level_nodes": (C.c_int, [_p, C.c_int, _i64p]),
    "hh_get_level_stencil": (C.c_int, [_p, C.c_int, _p]),
    "hh_get_diagonal": (C.c_int, [_p, C.c_int, C.c_double, _dp]),
    "hh_apply": (C.c_int, [_p, _p, _p, C.c_int64, C.c_int, C.c_double, C.c_int]),
    "hh_apply_device": (C.c_int, [_p, _p, _p, C.c_int64, C.c_int, C.c_double, C.c_int]),
    "hh_cycle": (C.c_int, [_p, _p, _p, C.c_int64]),
    "hh_cycle_device": (C.c_int, [_p, _p, _p, C.c_int64]),
    "hh_solve": (C.c_int, [_p, _p, _p, C.c_int64, C.POINTER(hh_solve_options), C.POINTER(C.c_int32), _dp]),
    "hh_solve_device": (C.c_int, [_p, _p, _p, C.c_int64, C.POINTER(hh_solve_options), C.POINTER(C.c_int32), _dp]),
    "hh_solve_point_sources": (C.c_int, [_p, _i64p, _dp, C.c_int64, _p, C.POINTER(hh_solve_options),
                                         C.POINTER(C.c_int32), _dp]),
    "hh_get_counters": (C.c_int, [_p, _dp, _dp, _i64p, _i64p]),
    "hh_profile_enable": (C.c_int, [_p, C.c_int]),
    "hh_profile_reset": (C.c_int, [_p]),
    "hh_profile_num_tags": (C.c_int, []),
    "hh_profile_tag_name": (C.c_char_p, [C.c_int]),
    "hh_profile_get": (C.c_int, [_p, C.c_int, _i64p, _dp, _dp]),
    "hh_profile_num_entries": (C.c_int, [_p]),
    "hh_profile_entry": (C.c_int, [_p, C.c_int, _ip, _i64p, _dp, _dp]),
}

_lib = None


def load():
    """Load the shared library (once) and attach signatures.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, handle=None):
    """Raise on hard errors (<0); return soft status (0 / HH_NOT_CONVERGED)."""
    if rc < 0:
        msg = load().hh_last_error(handle)
        raise HelmholtzB200Error(rc, msg.decode() if msg else "")
    return rc
