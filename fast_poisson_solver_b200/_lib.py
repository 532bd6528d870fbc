"""ctypes binding of libfps_sm100.so (C ABI declared in include/fps_sm100.h).

The library is built in-tree by ``fast_poisson_solver_b200.build`` (nvcc, sm_100a).  There is no
Python/CPU fallback: if the shared object is missing, ``load()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfps_sm100.so")

F = 700
FP = 704
FPW = 768  # weight / plane rows: 6 UMMA tiles of 128 features
LROWS = 1440  # FPS_LROWS: factor buffer rows (L + diagonal-block inverses + explicit inverse)
K0 = 400
NFREQ = 200
NHIDDEN = 5
ENGINE_SIMT = 0
ENGINE_TCGEN05 = 1
XOP_BYTES = 16384  # FPS_XOP_BYTES

_f32p = C.POINTER(C.c_float)


class FpsWeights(C.Structure):
    _fields_ = [
        ("h_freqs", _f32p),
        ("h_coefficients", _f32p),
        ("h_biases", _f32p),
        ("h_w0", _f32p),
        ("h_b0", _f32p),
        ("h_wh", _f32p * NHIDDEN),
        ("h_bh", _f32p * NHIDDEN),
    ]


# name -> (restype, argtypes); every symbol include/fps_sm100.h declares
_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_double
SIGNATURES = {
    "fps_last_error": (C.c_char_p, []),
    "fps_version": (_i32, []),
    "fps_create": (_i32, [C.POINTER(_vp), _i32, C.POINTER(FpsWeights), _i32, _i32]),
    "fps_destroy": (_i32, [_vp]),
    "fps_set_engine": (_i32, [_vp, _i32]),
    "fps_get_engine": (_i32, [_vp]),
    "fps_calibrate_scales": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "fps_get_scales": (_i32, [_vp, C.POINTER(C.c_float)]),
    "fps_set_calibration": (_i32, [_vp, C.POINTER(C.c_float), _i32]),
    "fps_get_calibration": (_i32, [_vp, C.POINTER(_f64)]),
    "fps_get_dh_calibration": (_i32, [_vp, C.POINTER(C.c_float)]),
    "fps_set_dh_calibration": (_i32, [_vp, C.POINTER(C.c_float)]),
    "fps_workspace_bytes": (C.c_size_t, [_vp]),
    "fps_features": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "fps_gram_accum": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fps_colsum_accum": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fps_assemble_factor": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(_f64), C.POINTER(_f64), _i32, _vp, _vp, _vp]),
    "fps_assemble_factor_packed": (_i32, [_vp, _vp, C.POINTER(_f64), C.POINTER(_f64), _i32, _vp, _vp, _vp]),
    "fps_project_rhs": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "fps_solve_factored": (_i32, [_vp, _vp, _vp, _i64, _i32, _f64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "fps_reconstruct": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i32, _vp, _i64, _vp]),
    "fps_sse_columns": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    "fps_sse_columns_f64": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    "fps_exact_min_rows": (_i64, []),
    "fps_gram_planes_bytes": (C.c_size_t, [_i64]),
    "fps_recon_planes_bytes": (C.c_size_t, [_i64]),
    "fps_reserve_workspace": (_i32, [_vp, _i64, _i32]),
    "fps_features_x": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, C.POINTER(_i32), _vp]),
    "fps_gram_accum_x": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp]),
    "fps_project_rhs_x": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "fps_recon_planes_build": (_i32, [_vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp]),
    "fps_reconstruct_x": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _i64, _vp]),
    "fps_features_f64": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "fps_gram_accum_f64": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fps_colsum_accum_f64": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fps_project_rhs_f64": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "fps_reconstruct_f64": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i32, _vp, _i64, _vp]),
    "fps_sine_sources": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _i64, _vp]),
    "fps_sine_sources_f64": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _i64, _vp]),
    "fps_perlin_sources": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _i64, _vp]),
    "fps_perlin_sources_f64": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _i64, _vp]),
    "fps_grid_points": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "fps_fd_workspace_bytes": (C.c_size_t, [_i32, _i32, _i32]),
    "fps_fd_solve": (_i32, [_vp, _i32, _i32, _f64, _f64, _vp, _i64, _vp, _i32, _vp, _i64, _vp, C.c_size_t, _vp]),
    "fps_fd_solve_f32": (_i32, [_vp, _i32, _i32, _f64, _f64, _vp, _i64, _vp, _i32, _vp, _i64, _vp, C.c_size_t, _vp]),
    "fps_comm_unique_id": (_i32, [_vp]),
    "fps_comm_init": (_i32, [C.POINTER(_vp), _i32, _vp, _i32, _i32]),
    "fps_allreduce_f64": (_i32, [_vp, _vp, _i64, _vp]),
    "fps_comm_destroy": (_i32, [_vp]),
    "fps_comm_abort": (_i32, [_vp]),
    "fps_launch_count": (_i64, []),
    "fps_profile_enable": (_i32, [_vp, _i32]),
    "fps_profile_reset": (_i32, [_vp]),
    "fps_profile_read": (_i32, [_vp, _i32, C.POINTER(_f64), C.POINTER(_i64)]),
    "fps_debug_tc_stats": (_i32, [_vp, _i32]),
}
PROF_KINDS = ["fourier", "layer0", "hidden", "last", "gram", "factor", "project", "trisolve", "reconstruct"]

_lib = None


class FpsError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FpsError(
            f"{LIB_PATH} is missing - build it with `python -m fast_poisson_solver_b200.build`. "
            "This package has no CPU or pure-PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().fps_last_error()
        raise FpsError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
