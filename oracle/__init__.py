"""TEST INFRASTRUCTURE ONLY: CPU oracles for the Demons hot path.

Two checkers live here, both behind the same ctypes interface (`OracleEngine`, `demons_run`):

* ``kind="reference"`` -- ``oracle/_ref/libdemons_ref.so``: the UNMODIFIED reference engine
  (`/root/reference/src/Alignment_Demons.cc:52-704`) compiled from its own sources by ``oracle/build_ref.py``.
* ``kind="port"``      -- ``oracle/_build/libdemons_oracle.so``: our plain-C restatement (``oracle/demons_oracle.c``),
  pinned bit-for-bit to the reference engine by ``tests/test_oracle.py`` and to ``tests/golden/*.npz``.
* ``kind="port_f32"``  -- the same C file built with fp32 field storage/accumulation: a CPU *prediction* of the drift of
  the GPU's fp32-storage mode, not a parity oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package.  Nothing under ``dicomautomaton_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


class Geom(C.Structure):
    _fields_ = [("slices", C.c_int64), ("rows", C.c_int64), ("cols", C.c_int64), ("channels", C.c_int64),
                ("pxl_dx", C.c_double), ("pxl_dy", C.c_double), ("pxl_dz", C.c_double)]


class Params(C.Structure):
    """Mirror of the loop-relevant fields of AlignViaDemonsParams (src/Alignment_Demons.h:20-61)."""
    _fields_ = [("max_iterations", C.c_int64), ("convergence_threshold", C.c_double),
                ("deformation_field_smoothing_sigma", C.c_double), ("update_field_smoothing_sigma", C.c_double),
                ("use_diffeomorphic", C.c_int32), ("_pad", C.c_int32),
                ("normalization_factor", C.c_double), ("max_update_magnitude", C.c_double)]


def make_params(max_iterations=100, convergence_threshold=0.001, deformation_field_smoothing_sigma=1.0,
                update_field_smoothing_sigma=0.5, use_diffeomorphic=False, normalization_factor=1.0,
                max_update_magnitude=2.0) -> Params:
    return Params(int(max_iterations), float(convergence_threshold), float(deformation_field_smoothing_sigma),
                  float(update_field_smoothing_sigma), 1 if use_diffeomorphic else 0, 0,
                  float(normalization_factor), float(max_update_magnitude))


def make_geom(shape, spacing=(1.0, 1.0, 1.0), channels=1) -> Geom:
    """shape = (slices, rows, cols); spacing = (pxl_dx, pxl_dy, pxl_dz) in mm."""
    s, r, c = (int(v) for v in shape)
    return Geom(s, r, c, int(channels), float(spacing[0]), float(spacing[1]), float(spacing[2]))


def build_port(force: bool = False, verbose: bool = False) -> dict:
    """Compile oracle/demons_oracle.c into the fp64 oracle and the fp32-storage emulation."""
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "demons_oracle.c")
    out = {}
    for name, defs in (("libdemons_oracle.so", []),
                       ("libdemons_oracle_f32.so", ["-DORC_FIELD_T=float", "-DORC_ACC_T=float", "-DORC_SUFFIX=_f32"])):
        lib = os.path.join(BUILD, name)
        if force or not os.path.isfile(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
            cmd = ["gcc", "-std=c11", "-O3", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", *defs, src,
                   "-o", lib, "-lm"]
            if verbose:
                print("[oracle]", " ".join(cmd))
            subprocess.check_call(cmd)
        out[name] = lib
    return out


def build_all(verbose: bool = False) -> dict:
    from . import build_ref
    libs = build_port(verbose=verbose)
    libs["libdemons_ref.so"] = build_ref.build(verbose=verbose)
    return libs


_LIBS: dict = {}


def _load(kind: str):
    if kind in _LIBS:
        return _LIBS[kind]
    if kind == "reference":
        path = os.path.join(HERE, "_ref", "libdemons_ref.so")
        if not os.path.isfile(path):
            from . import build_ref
            path = build_ref.build(verbose=False)
        if not path or not os.path.isfile(path):
            raise FileNotFoundError("oracle/_ref/libdemons_ref.so is absent and /root/reference is not available")
        prefix, suffix = "ref_", ""
    elif kind in ("port", "port_f32"):
        libs = build_port()
        path = libs["libdemons_oracle.so" if kind == "port" else "libdemons_oracle_f32.so"]
        prefix, suffix = "orc_", ("" if kind == "port" else "_f32")
    else:
        raise ValueError(kind)
    lib = C.CDLL(path)

    def fn(name, restype, argtypes):
        f = getattr(lib, prefix + name + suffix)
        f.restype, f.argtypes = restype, argtypes
        return f

    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)
    api = {
        "create": fn("engine_create", vp, [C.POINTER(Geom), vp, vp, C.POINTER(Params)]),
        "destroy": fn("engine_destroy", None, [vp]),
        "iterate": fn("engine_iterate", C.c_double, [vp]),
        "compute_update_and_mse": fn("engine_compute_update_and_mse", None, [vp, dp, ip]),
        "smooth_update": fn("engine_smooth_update", None, [vp, C.c_double]),
        "smooth_deformation": fn("engine_smooth_deformation", None, [vp, C.c_double]),
        "add_update": fn("engine_add_update", None, [vp]),
        "add_update_diffeomorphic": fn("engine_add_update_diffeomorphic", None, [vp]),
        "warp_moving": fn("engine_warp_moving", None, [vp]),
        "get": fn("engine_get", C.c_int, [vp, C.c_int, vp]),
        "set": fn("engine_set", C.c_int, [vp, C.c_int, vp]),
        "run": fn("demons_run", C.c_int, [C.POINTER(Geom), vp, vp, C.POINTER(Params), vp, vp, ip]),
    }
    if kind == "reference":
        f = lib.ref_worker_threads
        f.restype, f.argtypes = C.c_uint, []
        api["worker_threads"] = f
    _LIBS[kind] = api
    return api


def available(kind: str) -> bool:
    try:
        _load(kind)
        return True
    except (OSError, FileNotFoundError, subprocess.CalledProcessError):
        return False


def worker_threads(kind: str) -> int:
    if kind == "reference":
        return int(_load(kind)["worker_threads"]())
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


_BUFS = {"fixed": 0, "moving": 1, "warped": 2, "gradient": 3, "update": 4, "deformation": 5}


class OracleEngine:
    """Stage-level access to a CPU oracle engine (same stage names as the reference engine's methods)."""

    def __init__(self, fixed: np.ndarray, moving: np.ndarray, params: Params, spacing=(1.0, 1.0, 1.0),
                 kind: str = "port"):
        self.api = _load(kind)
        fixed = np.ascontiguousarray(fixed, dtype=np.float32)
        moving = np.ascontiguousarray(moving, dtype=np.float32)
        if fixed.ndim == 3:
            fixed, moving = fixed[..., None], moving[..., None]
        assert fixed.shape == moving.shape and fixed.ndim == 4
        self.shape = fixed.shape[:3]
        self.channels = fixed.shape[3]
        self.geom = make_geom(self.shape, spacing, self.channels)
        self.params = params
        self.h = self.api["create"](C.byref(self.geom), fixed.ctypes.data, moving.ctypes.data, C.byref(params))
        if not self.h:
            raise RuntimeError("oracle engine creation failed")

    def close(self):
        if self.h:
            self.api["destroy"](self.h)
            self.h = None

    __del__ = close

    def iterate(self) -> float:
        return float(self.api["iterate"](self.h))

    def compute_update_and_mse(self):
        mse, n = C.c_double(0), C.c_int64(0)
        self.api["compute_update_and_mse"](self.h, C.byref(mse), C.byref(n))
        return mse.value, n.value

    def smooth_update(self, sigma_mm):
        self.api["smooth_update"](self.h, float(sigma_mm))

    def smooth_deformation(self, sigma_mm):
        self.api["smooth_deformation"](self.h, float(sigma_mm))

    def add_update(self):
        self.api["add_update"](self.h)

    def add_update_diffeomorphic(self):
        self.api["add_update_diffeomorphic"](self.h)

    def warp_moving(self):
        self.api["warp_moving"](self.h)

    def get(self, name: str) -> np.ndarray:
        which = _BUFS[name]
        if which <= 2:
            out = np.empty(self.shape + (self.channels,), dtype=np.float32)
        else:
            out = np.empty(self.shape + (3,), dtype=np.float64)
        assert self.api["get"](self.h, which, out.ctypes.data) == 0
        return out

    def set(self, name: str, arr: np.ndarray):
        which = _BUFS[name]
        if which <= 2:
            arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(self.shape + (self.channels,))
        else:
            arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(self.shape + (3,))
        assert self.api["set"](self.h, which, arr.ctypes.data) == 0


def demons_run(fixed: np.ndarray, moving: np.ndarray, params: Params, spacing=(1.0, 1.0, 1.0), kind: str = "port"):
    """Whole registration loop (AlignViaDemons cc:977-997 on raw volumes).

    Returns (deformation[z,y,x,3] float64 mm, mse_history[iters], iterations_done)."""
    api = _load(kind)
    fixed = np.ascontiguousarray(fixed, dtype=np.float32)
    moving = np.ascontiguousarray(moving, dtype=np.float32)
    if fixed.ndim == 3:
        fixed, moving = fixed[..., None], moving[..., None]
    geom = make_geom(fixed.shape[:3], spacing, fixed.shape[3])
    D = np.empty(fixed.shape[:3] + (3,), dtype=np.float64)
    hist = np.full(max(int(params.max_iterations), 1), np.nan, dtype=np.float64)
    done = C.c_int64(0)
    rc = api["run"](C.byref(geom), fixed.ctypes.data, moving.ctypes.data, C.byref(params), D.ctypes.data,
                    hist.ctypes.data, C.byref(done))
    if rc != 0:
        raise RuntimeError("oracle demons_run failed")
    return D, hist[:done.value].copy(), int(done.value)
