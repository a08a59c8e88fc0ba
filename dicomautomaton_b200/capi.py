"""ctypes binding of the C ABI in include/dcma_b200/demons_c_api.h (the only way Python reaches the CUDA kernels).

There is no fallback of any kind here: if `lib/libdcma_b200.so` is missing the import of `lib()` raises, and every
compute entry point of the library itself fails with DCMA_B200_ENODEVICE when no sm_100a device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DCMA_B200_LIB selects another build of the SAME library (kernel-tuning experiments: tools/build_variant.py)
LIB_PATH = os.environ.get("DCMA_B200_LIB") or os.path.join(HERE, "lib", "libdcma_b200.so")

OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5
FIELD_FP64, FIELD_FP64_STRICT, FIELD_FP32, FIELD_MIXED = 0, 1, 2, 3
FIELD_MODES = {"fp64": FIELD_FP64, "fp64_strict": FIELD_FP64_STRICT, "fp32": FIELD_FP32, "mixed": FIELD_MIXED}
(STAGE_FORCE, STAGE_SMOOTH_U, STAGE_ADD, STAGE_COMPOSE, STAGE_SMOOTH_D, STAGE_WARP, STAGE_REDUCE,
 STAGE_GRADIENT) = range(8)
STAGE_NAMES = ["force", "smooth_update", "add_update", "compose", "smooth_deformation", "warp", "reduce", "gradient"]
BUF_FIXED, BUF_MOVING, BUF_WARPED, BUF_GRADIENT, BUF_UPDATE, BUF_DEFORMATION = range(6)


class Geom(C.Structure):
    _fields_ = [("slices", C.c_int64), ("rows", C.c_int64), ("cols", C.c_int64), ("channels", C.c_int64),
                ("pxl_dx", C.c_double), ("pxl_dy", C.c_double), ("pxl_dz", C.c_double)]


class Grid(C.Structure):
    """dcma_b200_grid: a regular grid in world space (geom + origin + unit axes)."""
    _fields_ = [("geom", Geom), ("origin", C.c_double * 3), ("ex", C.c_double * 3), ("ey", C.c_double * 3),
                ("ez", C.c_double * 3)]


class Params(C.Structure):
    _fields_ = [("max_iterations", C.c_int64), ("convergence_threshold", C.c_double),
                ("deformation_field_smoothing_sigma", C.c_double), ("update_field_smoothing_sigma", C.c_double),
                ("use_diffeomorphic", C.c_int32), ("use_histogram_matching", C.c_int32),
                ("histogram_bins", C.c_int64), ("histogram_outlier_fraction", C.c_double),
                ("normalization_factor", C.c_double), ("max_update_magnitude", C.c_double),
                ("verbosity", C.c_int64),
                ("field_mode", C.c_int32), ("device", C.c_int32), ("check_every", C.c_int32),
                ("pyramid_levels", C.c_int32), ("use_symmetric_force", C.c_int32), ("fused", C.c_int32),
                ("n_devices", C.c_int32), ("halo_planes", C.c_int32), ("pyramid_iterations", C.c_int32 * 3),
                ("reserved", C.c_int32 * 1)]


class RunStats(C.Structure):
    _fields_ = [("iterations_done", C.c_int64), ("kernel_launches", C.c_int64), ("loop_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("stage_ms", C.c_double * 8),
                ("stage_launches", C.c_int64 * 8), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


# every symbol the header declares, with its signature: (restype, argtypes)
_vp, _dp, _ip, _cp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_char_p
SYMBOLS = {
    "dcma_b200_abi_version": (C.c_int, []),
    "dcma_b200_last_error": (C.c_char_p, []),
    "dcma_b200_device_count": (C.c_int, []),
    "dcma_b200_demons_params_default": (None, [C.POINTER(Params)]),
    "dcma_b200_demons_register": (C.c_int, [C.POINTER(Geom), _vp, _vp, C.POINTER(Params), _vp, _vp,
                                            C.POINTER(RunStats), _cp, C.c_size_t]),
    "dcma_b200_engine_create": (C.c_int, [C.POINTER(_vp), C.POINTER(Geom), C.POINTER(Params), _vp, _cp, C.c_size_t]),
    "dcma_b200_engine_destroy": (None, [_vp]),
    "dcma_b200_engine_load": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "dcma_b200_engine_run": (C.c_int, [_vp, C.c_int64, C.c_int, _vp, C.POINTER(RunStats)]),
    "dcma_b200_engine_stage": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _ip]),
    "dcma_b200_engine_get": (C.c_int, [_vp, C.c_int, _vp]),
    "dcma_b200_engine_set": (C.c_int, [_vp, C.c_int, _vp]),
    "dcma_b200_engine_export_deformation_device": (C.c_int, [_vp, _vp]),
    "dcma_b200_engine_sync": (C.c_int, [_vp]),
    "dcma_b200_engine_device_bytes": (C.c_int64, [_vp]),
    "dcma_b200_warp_image": (C.c_int, [C.POINTER(Geom), _vp, _vp, C.c_float, C.c_int, _vp, C.c_int, _cp, C.c_size_t]),
    "dcma_b200_restrict_image": (C.c_int, [C.POINTER(Geom), _vp, C.POINTER(Geom), _vp, C.c_int, _cp, C.c_size_t]),
    "dcma_b200_prolong_field": (C.c_int, [C.POINTER(Geom), _vp, _vp, C.c_int, _cp, C.c_size_t]),
    "dcma_b200_histogram_match": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_double, _vp, C.POINTER(C.c_int32),
                                            C.c_int, _cp, C.c_size_t]),
    "dcma_b200_warp_image_grid": (C.c_int, [C.POINTER(Grid), _vp, C.POINTER(Grid), _vp, _vp, C.c_float, _vp, C.c_int, _cp,
                                            C.c_size_t]),
    "dcma_b200_warp_to_grid": (C.c_int, [C.POINTER(Grid), _vp, C.POINTER(Grid), _vp, C.POINTER(Grid), _vp, C.c_int32, C.c_float, _vp,
                                         C.c_int, _cp, C.c_size_t]),
    "dcma_b200_demons_register_stacks": (C.c_int, [C.POINTER(Grid), _vp, C.POINTER(Grid), _vp, C.POINTER(Params), _vp, _vp,
                                                   C.POINTER(RunStats), _vp, _vp, _cp, C.c_size_t]),
    "dcma_b200_device_pool_trim": (None, []),
    "dcma_b200_device_pool_bytes": (C.c_int64, []),
    "dcma_b200_virtual_field": (C.c_int, [C.c_int32, C.POINTER(Grid), _vp, _vp, _vp, C.c_int, _cp, C.c_size_t]),
    "dcma_b200_resample_to_grid": (C.c_int, [C.POINTER(Grid), _vp, C.POINTER(Grid), C.c_float, _vp, C.c_int, _cp, C.c_size_t]),
    "dcma_b200_format_doubles": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, C.POINTER(C.c_int64), C.c_int, _cp, C.c_size_t]),
    "dcma_b200_parse_doubles": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.POINTER(C.c_int64), C.c_int, _cp, C.c_size_t]),
    "dcma_b200_slab_partition": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int64, _ip, _ip, _ip, _ip]),
    "dcma_b200_slab_group_create_local": (C.c_int, [C.POINTER(_vp), C.POINTER(Geom), C.POINTER(Params),
                                                    C.POINTER(C.c_int32), C.c_int32, C.c_int64, _cp, C.c_size_t]),
    "dcma_b200_nccl_unique_id": (C.c_int, [_vp, _cp, C.c_size_t]),
    "dcma_b200_slab_group_create_nccl": (C.c_int, [C.POINTER(_vp), C.POINTER(Geom), C.POINTER(Params), _vp, C.c_int32,
                                                   C.c_int32, C.c_int32, C.c_int64, _cp, C.c_size_t]),
    "dcma_b200_slab_group_destroy": (None, [_vp]),
    "dcma_b200_slab_group_load": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "dcma_b200_slab_group_run": (C.c_int, [_vp, C.c_int64, _vp, C.POINTER(RunStats)]),
    "dcma_b200_slab_group_range": (C.c_int, [_vp, _ip, _ip]),
    "dcma_b200_slab_group_get_deformation": (C.c_int, [_vp, _vp]),
    "dcma_b200_host_alloc": (C.c_int, [_vp, C.c_size_t, _vp, C.c_size_t]),
    "dcma_b200_host_free": (None, [_vp]),
    "dcma_b200_slab_group_halo_bytes": (C.c_int64, [_vp]),
    "dcma_b200_slab_group_exchange_stats": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
}

_LIB = None


class DcmaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"dcma_b200 error {code}: {message}")
        self.code = code


def lib():
    """Load the C-ABI library; raises if it has not been built (run `python -m dicomautomaton_b200.build`)."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback); "
                              "build it with `python -m dicomautomaton_b200.build`")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _LIB = handle
    return _LIB


def check(rc: int, errbuf=None):
    if rc != OK:
        msg = errbuf.value.decode(errors="replace") if errbuf is not None and errbuf.value else \
            (lib().dcma_b200_last_error() or b"").decode(errors="replace")
        raise DcmaError(rc, msg)


def default_params() -> Params:
    p = Params()
    lib().dcma_b200_demons_params_default(C.byref(p))
    return p
