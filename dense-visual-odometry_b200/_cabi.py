"""ctypes binding of libdvo_b200.so (include/dvo_b200.h).  No CPU fallback: if the library is missing
or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

DVO_MAX_LEVELS = 8
DVO_ACC_TERMS = 29
W_NONE, W_TDIST_REF, W_HUBER, W_HUBER_MAD = 0, 1, 2, 3
OOB_INCLUSIVE, OOB_STRICT = 0, 1

# developer knob: DVO_B200_LIB points the binding at another build of the same C ABI (kernel experiments)
LIB_PATH = Path(os.environ.get("DVO_B200_LIB") or (Path(__file__).resolve().parent / "libdvo_b200.so"))


class DvoError(RuntimeError):
    pass


class dvo_config(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int32), ("max_increased_steps", C.c_int32), ("tolerance", C.c_float),
        ("sigma_prior", C.c_float), ("weights", C.c_int32), ("oob_mode", C.c_int32), ("tdist_dof", C.c_float),
        ("tdist_init_sigma", C.c_float), ("tdist_tolerance", C.c_float), ("tdist_max_iterations", C.c_int32),
        ("huber_k", C.c_float), ("max_distance", C.c_float), ("threads_per_block", C.c_int32),
        ("blocks_per_sm", C.c_int32), ("approximate_image2_gradient", C.c_int32), ("cluster_size", C.c_int32), ("tdist_mean", C.c_int32),
        ("use_depth_residual", C.c_int32), ("depth_weight", C.c_float), ("reserved", C.c_int32 * 1),
    ]


class dvo_pair_stats(C.Structure):
    _fields_ = [
        ("iters", C.c_int32 * DVO_MAX_LEVELS), ("n_valid", C.c_int32 * DVO_MAX_LEVELS),
        ("err", C.c_float * DVO_MAX_LEVELS), ("flags", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


STATS_BYTES = C.sizeof(dvo_pair_stats)
assert STATS_BYTES == 128

# every symbol include/dvo_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "dvo_default_config": (None, [C.POINTER(dvo_config)]),
    "dvo_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.POINTER(dvo_config)]),
    "dvo_destroy": (C.c_int, [_P]),
    "dvo_last_error": (C.c_char_p, [_P]),
    "dvo_set_intrinsics": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double]),
    "dvo_build_pyramids": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "dvo_build_pyramids_gray": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "dvo_build_pyramids_host": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "dvo_upload_frames": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "dvo_build_pyramids_staged": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "dvo_depth_clamp_threshold": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "dvo_estimate": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "dvo_estimate_host": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "dvo_residuals_jacobian": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "dvo_depth_residuals_jacobian": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "dvo_get_pyramid": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "dvo_get_point_list": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "dvo_level_shape": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dvo_level_intrinsics": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "dvo_launch_count": (C.c_longlong, [_P]),
    "dvo_last_estimate_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "dvo_debug_bounds_violations": (C.c_int, [_P, C.POINTER(C.c_ulonglong)]),
}

_lib = None


def load():
    """Loads the shared library once and declares all prototypes."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DvoError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                           " (there is no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(lib, handle, rc, what):
    if rc != 0:
        msg = lib.dvo_last_error(handle)
        raise DvoError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
