"""ctypes binding of libphyclone_b200.so (the C ABI in include/phyclone_b200.h)."""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libphyclone_b200.so")


class PcbError(RuntimeError):
    """Raised for every non-zero pcb_status; carries pcb_last_error()."""


class Layout(C.Structure):
    _fields_ = [
        ("S", C.c_int32),
        ("G", C.c_int32),
        ("strip", C.c_int32),
        ("lanes_per_row", C.c_int32),
        ("n_strips", C.c_int32),
        ("pitch", C.c_int32),
        ("lead", C.c_int32),
        ("rows_per_cta", C.c_int32),
        ("kernel", C.c_int32),
        ("reserved", C.c_int32),
        ("tile_elems", C.c_int64),
    ]


class Arena(C.Structure):
    _fields_ = [
        ("layout", Layout),
        ("lg", C.c_void_p),
        ("lin", C.c_void_p),
        ("rmax", C.c_void_p),
        ("capacity", C.c_int64),
    ]


_P = C.c_void_p
_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol include/phyclone_b200.h declares
SIGNATURES = {
    "pcb_version": (C.c_int, []),
    "pcb_last_error": (C.c_char_p, []),
    "pcb_device_count": (C.c_int, []),
    "pcb_set_device": (C.c_int, [C.c_int]),
    "pcb_launch_count": (_i64, []),
    "pcb_make_layout": (C.c_int, [_i32, _i32, _i32, C.POINTER(Layout)]),
    "pcb_arena_import_dev": (C.c_int, [C.POINTER(Arena), _P, _P, _i64, _P]),
    "pcb_arena_export_dev": (C.c_int, [C.POINTER(Arena), _P, _i64, _P, _P]),
    "pcb_tile_add_dev": (C.c_int, [C.POINTER(Arena), _P, _P, _P, _i64, _f64, _f64, _P]),
    "pcb_node_jobs_dev": (C.c_int, [C.POINTER(Arena), _P, _P, _i64, _f64, _P, _P, _P]),
    "pcb_run_plan_dev": (C.c_int, [C.POINTER(Arena), _P, _i64, _P, _P, _i32, _f64, _P, _P, _i64, _P, _P, _P, _P]),
    "pcb_plan_stats": (C.c_int, [C.POINTER(_f64), C.POINTER(_f64), C.POINTER(_i64)]),
    "pcb_timing_begin": (C.c_int, [_i32]),
    "pcb_timing_end": (C.c_int, [C.POINTER(_f64), C.POINTER(_i64)]),
    "pcb_timing_sample": (C.c_int, [_i32, C.POINTER(_i64)]),
    "pcb_reduce_rows_dev": (C.c_int, [_P, _i64, _i32, _P, _P]),
    "pcb_np_conv_dims_host": (C.c_int, [_P, _P, _P, _i64, _i32, _i32]),
    "pcb_sub_compute_S_host": (C.c_int, [_P, _P, _i64, _i32, _i32]),
    "pcb_compute_log_S_host": (C.c_int, [_P, _P, _P, _i64, _i32, _i32]),
    "pcb_node_update_host": (C.c_int, [_P, _P, _P, _P, _i64, _i32, _i32]),
    "pcb_root_terms_host": (C.c_int, [_P, _P, _P, _i64, _i32, _i32]),
    "pcb_likelihood_grid_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _i64, _i32, _i32, _i32, _i32, _f64, _P]),
    "pcb_likelihood_grid_dev": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _i64, _i32, _i32, _i32, _i32, _f64, _P, _P]),
    "pcb_cluster_sum_host": (C.c_int, [_P, _P, _i64, _i32, _i32, _P]),
    "pcb_outlier_marginal_host": (C.c_int, [_P, _i64, _i32, _i32, _P]),
    "pcb_map_trees_host": (C.c_int, [_P, _P, _P, _P, _P, _i64, _i64, _i64, _i32, _i32, _P, _P]),
    "pcb_map_trees_dev": (C.c_int, [_P, _P, _P, _P, _P, _i64, _i64, _i64, _i32, _i32, _P, _P, _P, _P, _P]),
    "pcb_conv_log_host": (C.c_int, [_P, _P, _P, _i64, _i32]),
    "pcb_swarm_weights_host": (C.c_int, [_P, _i64, _P, _P, _P, _P]),
    "pcb_normalize_segments_host": (C.c_int, [_P, _P, _i64, _P, _P, _P]),
    "pcb_normalize_segments_dev": (C.c_int, [_P, _P, _i64, _P, _P, _P, _P, _P]),
    "pcb_resample_host": (C.c_int, [_P, _i64, _P, _i64, _P]),
    "pcb_smc_sweep_dev": (C.c_int, [C.POINTER(Arena), _P, _P]),
    "pcb_fp64_peak": (C.c_int, [_i32, C.POINTER(_f64), C.POINTER(_f64)]),
}

_lib = None


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcbError(
                "%s is missing: build it with `python -m phyclone_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise PcbError("phyclone_b200 error %d: %s" % (rc, load().pcb_last_error().decode()))


def require_device():
    lib = load()
    if lib.pcb_device_count() < 1:
        raise PcbError("no CUDA device visible: phyclone_b200 has no CPU fallback")
    return lib
