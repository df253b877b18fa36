"""ctypes binding of libnkb200.so (the C ABI declared in include/nkb200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  The library is built in-tree by ``__graft_entry__.build()`` /
``newton-krylov_ooc_b200/csrc/Makefile``.
"""

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_size_t, c_uint64, c_void_p

NKB_MAX_TRACERS = 6
NKB_MAX_CLASSES = 3

MOD_LINEAR = 0
MOD_FORCED_FILE = 1
MOD_PHOSPHORUS = 2
MOD_PHOSPHORUS_1D = 3

# NKB_LIB_PATH selects another build of the SAME library (kernel experiments: scripts/ab_variants.sh); default in-tree
LIB_PATH = os.environ.get("NKB_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnkb200.so")


class NkbError(RuntimeError):
    """a call into libnkb200.so failed"""


class ModelDesc(ctypes.Structure):
    """mirror of struct nkb_model_desc (include/nkb200.h)"""

    _fields_ = [
        ("nz", c_int32),
        ("ny", c_int32),
        ("n_tracers", c_int32),
        ("kind", c_int32),
        ("n_classes", c_int32),
        ("class_of", c_int32 * NKB_MAX_TRACERS),
        ("column_model", c_int32),
        ("t0", c_double),
        ("t1", c_double),
        ("h_depth_edges", POINTER(c_double)),
        ("h_ypos_mid", POINTER(c_double)),
        ("h_wvel", POINTER(c_double)),
        ("h_estencil", POINTER(c_double)),
        ("h_bld_max", POINTER(c_double)),
        ("surf_diag", c_double * NKB_MAX_CLASSES),
        ("surf_aff", c_double * NKB_MAX_CLASSES),
        ("decay", c_double * NKB_MAX_CLASSES),
        ("sink_vel", c_double * NKB_MAX_CLASSES),
        ("n_flux_pts", c_int32),
        ("flux_t", c_double * 8),
        ("flux_v", c_double * 8),
        ("src_const", c_double * NKB_MAX_TRACERS),
        ("sink_thres", c_double),
        ("n_frc", c_int32),
        ("h_frc_time", POINTER(c_double)),
        ("h_frc_data", POINTER(c_double)),
        ("h_light", POINTER(c_double)),
        ("po4_halfsat", c_double),
        ("max_uptake_rate", c_double),
        ("sigma", c_double),
        ("dop_remin_rate", c_double),
        ("pop_remin_rate", c_double),
        ("po4_s_restoring_opt", c_int32),
        ("n_srf", c_int32),
        ("h_srf_time", POINTER(c_double)),
        ("h_srf_data", POINTER(c_double)),
        ("srf_rate", c_double * NKB_MAX_CLASSES),
    ]


# name -> (restype, argtypes); every symbol declared in include/nkb200.h
SYMBOLS = {
    "nkb_last_error": (c_char_p, []),
    "nkb_version": (c_int, []),
    "nkb_launch_count": (c_uint64, []),
    "nkb_model_create": (c_int, [POINTER(c_void_p), POINTER(ModelDesc)]),
    "nkb_model_destroy": (None, [c_void_p]),
    "nkb_model_set_schedule": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_double)]),
    "nkb_model_mixing_coeff": (c_int, [c_void_p, c_double, c_void_p, c_void_p]),
    "nkb_model_tend": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "nkb_model_work_doubles": (c_size_t, [c_void_p, c_int, c_int]),
    "nkb_model_eval": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_int), c_void_p, c_void_p],
    ),
    "nkb_model_eval_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "nkb_model_poll_error": (c_int, [c_void_p]),
    "nkb_banded_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, POINTER(c_double)]),
    "nkb_banded_destroy": (None, [c_void_p]),
    "nkb_banded_blocks": (c_int, [c_void_p]),
    "nkb_banded_path": (c_int, [c_void_p]),
    "nkb_banded_solve": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_void_p]),
    "nkb_pack_members": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nkb_unpack_members": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nkb_wdot_chunks": (c_int, [c_int]),
    "nkb_wdot": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
         c_int, c_void_p, c_void_p],
    ),
    "nkb_axpby": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int,
         c_int, c_void_p],
    ),
    "nkb_mgs_scratch_doubles": (c_size_t, [c_int, c_int, c_int]),
    "nkb_mgs": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, POINTER(c_void_p),
         c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p],
    ),
    "nkb_lin_comb": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_void_p, POINTER(c_void_p), c_int, c_void_p, c_void_p, c_double, c_int, c_int,
         c_void_p],
    ),
    "nkb_interleave_blocks": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "nkb_fd_sigma": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "nkb_limiter_scalef": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_double, c_int, c_double, c_int, c_int, c_int,
         c_void_p, c_void_p, c_void_p],
    ),
}

_lib = None


def load():
    """load libnkb200.so (once); raises NkbError when it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NkbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().nkb_last_error()
        raise NkbError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def dptr(arr):
    """POINTER(c_double) of a C-contiguous float64 numpy array (kept alive by the caller)"""
    return arr.ctypes.data_as(POINTER(c_double))
