"""
ctypes binding of libxcb200.so (the C ABI declared in include/xcb200.h).

There is NO fallback: if the shared library is missing, or no CUDA device is
visible when a compute entry point is called, the call raises.  The library is
built in-tree by ``python -m xcontour_b200.build`` (nvcc, sm_100a).
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_int, c_int32,
                    c_long, c_size_t, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XCB200_LIB") or os.path.join(HERE, "libxcb200.so")   # XCB200_LIB: A/B builds only

XC_F32, XC_F64, XC_F32_AS_F64 = 0, 1, 2
SCAN_PREFIX, SCAN_TOTAL_MINUS, SCAN_SUFFIX = 0, 1, 2
PART = {"all": 0, "upper": 1, "lower": 2}
MAX_INTEGRANDS = 3
N_STAGES = 5
STAGE_NAMES = ("minmax_levels", "edges", "bin_accumulate", "epilogue", "lwa")


class KeffLwaArgs(Structure):
    """Mirror of ``xc_keff_lwa_args`` (include/xcb200.h)."""
    _fields_ = [
        ("q", c_void_p), ("q_dtype", c_int),
        ("S", c_long), ("n_y", c_int), ("n_x", c_int),
        ("N", c_int), ("increase", c_int), ("lt", c_int),
        ("ctr_dtype", c_int),
        ("dA", c_void_p), ("dA_dtype", c_int),
        ("grdS", c_void_p), ("grdS_dtype", c_int),
        ("lat_rad", c_void_p), ("dlambda", c_double),
        ("table", c_void_p), ("table_coord", c_void_p), ("n_table", c_int),
        ("eq_coord", c_void_p),
        ("ww", c_void_p),
        ("keff_mask", c_double),
        ("part", c_int), ("sub_batch", c_int),
        ("ctr", c_void_p), ("area", c_void_p), ("intgrdS", c_void_p),
        ("latEq", c_void_p), ("Lmin", c_void_p), ("dintSdA", c_void_p),
        ("dqdA", c_void_p), ("Leq2", c_void_p), ("nkeff", c_void_p),
        ("Qref", c_void_p), ("lwa", c_void_p),
        ("stage_ms", c_void_p),
        ("cx", c_void_p), ("cy", c_void_p), ("bcx", c_int), ("bcy", c_int), ("fill_value", c_double),
        ("dA_row", c_void_p), ("uniform_dA", c_int), ("any_degenerate", c_int),
        ("ww_row", c_void_p), ("lwa_f32", c_int),
        ("numpy2_rules", c_int),
    ]


# name -> (restype, argtypes); every symbol of include/xcb200.h
SIGNATURES = {
    "xc_last_error": (c_char_p, []),
    "xc_abi_version": (c_int, []),
    "xc_minmax_levels_workspace_bytes": (c_size_t, [c_long, c_long]),
    "xc_minmax_levels": (c_int, [c_void_p, c_int, c_long, c_long, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xc_hist_edges": (c_int, [c_void_p, c_long, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "xc_equal_area_levels_workspace_bytes": (c_size_t, [c_long, c_long, c_int, c_int]),
    "xc_equal_area_levels": (c_int, [c_void_p, c_int, c_long, c_long, c_void_p, c_int,
                                     c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "xc_bin_accumulate_workspace_bytes": (c_size_t, [c_long, c_long, c_int, c_int]),
    "xc_bin_accumulate": (c_int, [c_void_p, c_int, c_long, c_long,
                                  c_void_p, c_long, c_int, c_int,
                                  c_void_p, c_int, c_int,
                                  POINTER(c_void_p), POINTER(c_int), c_int,
                                  c_void_p, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "xc_interp": (c_int, [c_void_p, c_long, c_int, c_void_p, c_long, c_void_p, c_long,
                          c_int, c_int, c_long, c_void_p, c_void_p]),
    "xc_gradient_wrt_area": (c_int, [c_void_p, c_int, c_void_p, c_int, c_long, c_int,
                                     c_void_p, c_void_p]),
    "xc_gradient_wrt_area_coord": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                           c_long, c_int, c_void_p, c_void_p]),
    "xc_leq2": (c_int, [c_void_p, c_void_p, c_long, c_void_p, c_void_p]),
    "xc_lmin": (c_int, [c_void_p, c_long, c_void_p, c_void_p]),
    "xc_nkeff": (c_int, [c_void_p, c_void_p, c_double, c_long, c_void_p, c_void_p]),
    "xc_eqlat": (c_int, [c_void_p, c_long, c_void_p, c_void_p]),
    "xc_lwa_weights_workspace_bytes": (c_size_t, [c_long]),
    "xc_lwa_weights": (c_int, [c_void_p, c_int, c_long, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xc_lwa_workspace_bytes": (c_size_t, [c_long]),
    "xc_lwa": (c_int, [c_void_p, c_int, c_long, c_int, c_int, c_void_p, c_void_p,
                       c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xc_lwa_ex": (c_int, [c_void_p, c_int, c_long, c_int, c_int, c_void_p, c_void_p, c_void_p,
                          c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xc_lwa_mask": (c_int, [c_void_p, c_int, c_long, c_int, c_int, c_void_p, c_int, c_int, c_int,
                            c_void_p, c_void_p]),
    "xc_grad2_latlon": (c_int, [c_void_p, c_int, c_long, c_int, c_int, c_void_p, c_double,
                                c_void_p, c_int, c_void_p]),
    "xc_latlon_cell_area": (c_int, [c_void_p, c_int, c_int, c_double, c_void_p, c_int, c_void_p]),
    "xc_keff_lwa_batch_workspace_bytes": (c_size_t, [c_long, c_int, c_int, c_int]),
    "xc_keff_lwa_batch": (c_int, [POINTER(KeffLwaArgs), c_void_p, c_size_t, c_void_p]),
    "xc_launch_count": (c_long, []),
    "xc_reset_launch_count": (None, []),
}

_lib = None


def load():
    """Load libxcb200.so and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "xcontour_b200: %s is missing -- build it with "
            "`python -m xcontour_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Turn a non-zero status into the reference's error convention: a bare
    ``Exception`` carrying the message (xcontour/core.py:54, 57, 733, ...)."""
    if rc != 0:
        raise Exception(load().xc_last_error().decode())
