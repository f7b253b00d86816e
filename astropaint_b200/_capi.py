"""ctypes binding of libastropaint_b200.so (include/astropaint_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing, or a call fails, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AP_LIB_PATH") or os.path.join(_HERE, "_native", "libastropaint_b200.so")

_lib = None

c_i64 = C.c_int64
c_i32 = C.c_int32
c_dbl = C.c_double
c_ptr = C.c_void_p


class ApHalos(C.Structure):
    _fields_ = [("n", c_i64), ("d_x", c_ptr), ("d_y", c_ptr), ("d_z", c_ptr), ("d_radius", c_ptr),
                ("d_theta", c_ptr), ("d_phi", c_ptr), ("d_D_a", c_ptr)]


#: every symbol include/astropaint_b200.h declares -> (restype, argtypes)
SIGNATURES = {
    "ap_last_error": (C.c_char_p, []),
    "ap_version": (C.c_int, []),
    "ap_debug_bounds_errors": (C.c_longlong, []),
    "ap_disc_mode": (C.c_int, [C.c_int]),
    "ap_quad_fast_path": (C.c_int, [C.c_int]),
    "ap_catalog_build": (C.c_int, [c_i64] + [c_ptr] * 8 + [c_dbl, c_dbl, c_dbl, C.POINTER(c_ptr), c_ptr]),
    "ap_disc_radius": (C.c_int, [c_i64, c_ptr, c_dbl, c_ptr, c_ptr]),
    "ap_disc_count": (C.c_int, [c_i64, C.POINTER(ApHalos), C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_band_clip": (C.c_int, [c_i64, c_i64]),
    "ap_disc_count_items": (C.c_int, [c_i64, C.POINTER(ApHalos), c_ptr, c_ptr, c_ptr, c_i32, c_ptr, c_i64, c_ptr, c_ptr]),
    "ap_paint_los_items": (C.c_int, [C.c_int, c_i64, C.POINTER(ApHalos), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr,
                                     c_ptr, c_ptr, c_ptr]),
    "ap_paint_los_items_f32": (C.c_int, [C.c_int, c_i64, C.POINTER(ApHalos), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                                         c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_emit_begin": (C.c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i32]),
    "ap_emit_end": (C.c_int, []),
    "ap_segment_accumulate": (C.c_int, [c_i64, c_ptr, c_ptr, c_i32, c_i64, c_ptr, C.c_int, c_ptr]),
    "ap_disc_near_ties": (C.c_int, [c_i64, C.POINTER(ApHalos), c_dbl, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "ap_query_disc_host": (C.c_int, [c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr, c_ptr, c_i64]),
    "ap_disc_pixels": (C.c_int, [c_i64, C.POINTER(ApHalos), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_scatter_add": (C.c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_paint_profile": (C.c_int, [C.c_int, c_i64, C.POINTER(ApHalos), C.POINTER(c_ptr), C.POINTER(c_i32),
                                   c_i32, c_ptr, c_ptr, c_ptr]),
    "ap_table_nodes": (C.c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    "ap_los_nodes": (C.c_int, [C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_b16_tau_eval": (C.c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_spline_build": (C.c_int, [c_i64, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_spline_build_uniform": (C.c_int, [c_i64, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_paint_table": (C.c_int, [c_i64, C.POINTER(ApHalos), c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                 c_ptr]),
    "ap_scatter_add_f32": (C.c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_paint_profile_f32": (C.c_int, [C.c_int, c_i64, C.POINTER(ApHalos), C.POINTER(c_ptr), C.POINTER(c_i32),
                                       c_i32, c_ptr, c_ptr, c_ptr]),
    "ap_paint_table_f32": (C.c_int, [c_i64, C.POINTER(ApHalos), c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                     c_ptr]),
    "ap_cutouts": (C.c_int, [c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl, c_dbl, c_dbl, c_dbl, c_i32, c_i32,
                             c_ptr, c_ptr]),
    "ap_stack_cutouts": (C.c_int, [c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl, c_dbl, c_dbl, c_dbl, c_i32,
                                   c_i32, c_ptr, c_ptr, c_ptr]),
    "ap_patch_spline_prefilter": (C.c_int, [c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "ap_patch_rotate": (C.c_int, [c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr]),
    "ap_patch_scale": (C.c_int, [c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_ud_grade": (C.c_int, [c_i64, c_ptr, c_i64, C.c_int, c_dbl, C.c_int, c_ptr, c_ptr]),
    "ap_ang2pix": (C.c_int, [c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_pix2ang": (C.c_int, [c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ap_spray_table_host": (C.c_int, [C.c_int, c_i64, c_i64] + [c_ptr] * 11 + [c_i32, c_i32, c_ptr, C.POINTER(C.c_uint64)]),
    "ap_host_register": (C.c_int, [c_ptr, C.c_uint64]),
    "ap_host_unregister": (C.c_int, [c_ptr]),
    "ap_copy_d2h_async": (C.c_int, [c_ptr, c_ptr, C.c_uint64, c_ptr]),
    "ap_copy_h2d_async": (C.c_int, [c_ptr, c_ptr, C.c_uint64, c_ptr]),
    "ap_fp64_peak": (C.c_int, [c_i32, c_i32, c_ptr, C.POINTER(C.c_uint64), c_ptr]),
    "ap_spray_host": (C.c_int, [C.c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                C.POINTER(c_ptr), C.POINTER(c_i32), c_i32, c_ptr, C.POINTER(C.c_uint64)]),
}


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "astropaint_b200 has no CPU fallback.")
    handle = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return _lib


class NativeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = lib().ap_last_error()
        raise NativeError(msg.decode() if msg else f"native call failed with code {rc}")


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return c_ptr(t.data_ptr())
    return c_ptr(t.ctypes.data)
