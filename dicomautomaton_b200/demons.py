"""Host-side mirror of the reference's Demons interface, driving the CUDA library through its C ABI.

Reference names kept on purpose (src/Alignment_Demons.h): `AlignViaDemonsParams` (:20-61), `demons_volume` (:64-74),
`AlignViaDemons` (:233-236).  The C++ drop-in for the dispatcher lives in `dicomautomaton_b200/host/`; this module is
what the Python tests and bench.py call.  Nothing here computes: every array operation happens in
`lib/libdcma_b200.so` on the GPU, and a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import sys
from typing import Optional

import numpy as np

from . import capi


@dataclasses.dataclass
class AlignViaDemonsParams:
    """Same fields and defaults as the reference struct (src/Alignment_Demons.h:20-61) + additive extensions."""
    max_iterations: int = 100
    convergence_threshold: float = 0.001
    deformation_field_smoothing_sigma: float = 1.0
    update_field_smoothing_sigma: float = 0.5
    use_diffeomorphic: bool = False
    use_histogram_matching: bool = False
    histogram_bins: int = 256
    histogram_outlier_fraction: float = 0.01
    normalization_factor: float = 1.0
    max_update_magnitude: float = 2.0
    verbosity: int = 1
    # extensions (defaults keep reference behaviour)
    field_mode: str = "fp64"  # "fp64" | "fp64_strict" | "fp32" | "mixed" (D fp64, U fp32)
    device: int = -1
    check_every: int = 0
    pyramid_levels: int = 1
    use_symmetric_force: bool = False
    fused: bool = True
    pyramid_iterations: tuple = (0, 0, 0)  # iterations at pyramid levels 1..3 (0 = max_iterations)
    n_devices: int = 0   # AlignViaDemons only: > 1 shards the volume into z-slabs over devices 0..n-1 of this process
    halo_planes: int = 0

    def to_c(self) -> capi.Params:
        p = capi.default_params()
        p.max_iterations = int(self.max_iterations)
        p.convergence_threshold = float(self.convergence_threshold)
        p.deformation_field_smoothing_sigma = float(self.deformation_field_smoothing_sigma)
        p.update_field_smoothing_sigma = float(self.update_field_smoothing_sigma)
        p.use_diffeomorphic = 1 if self.use_diffeomorphic else 0
        p.use_histogram_matching = 1 if self.use_histogram_matching else 0
        p.histogram_bins = int(self.histogram_bins)
        p.histogram_outlier_fraction = float(self.histogram_outlier_fraction)
        p.normalization_factor = float(self.normalization_factor)
        p.max_update_magnitude = float(self.max_update_magnitude)
        p.verbosity = int(self.verbosity)
        p.field_mode = capi.FIELD_MODES[self.field_mode]
        p.device = int(self.device)
        p.check_every = int(self.check_every)
        p.pyramid_levels = int(self.pyramid_levels)
        p.use_symmetric_force = 1 if self.use_symmetric_force else 0
        p.fused = 1 if self.fused else 0
        p.n_devices = int(self.n_devices)
        p.halo_planes = int(self.halo_planes)
        for i in range(3):
            p.pyramid_iterations[i] = int(self.pyramid_iterations[i]) if i < len(self.pyramid_iterations) else 0
        return p


@dataclasses.dataclass
class demons_volume:
    """Dense proxy of an image stack: reference `demons_volume<T>` (src/Alignment_Demons.h:64-74).

    data has shape (slices, rows, cols, channels): idx = ((z*rows + y)*cols + x)*channels + c (:77-79)."""
    data: np.ndarray
    pxl_dx: float = 1.0
    pxl_dy: float = 1.0
    pxl_dz: float = 1.0

    def __post_init__(self):
        if self.data.ndim == 3:
            self.data = self.data[..., None]
        if self.data.ndim != 4:
            raise ValueError("demons_volume data must be (slices, rows, cols[, channels])")

    @property
    def slices(self): return self.data.shape[0]
    @property
    def rows(self): return self.data.shape[1]
    @property
    def cols(self): return self.data.shape[2]
    @property
    def channels(self): return self.data.shape[3]

    def geom(self) -> capi.Geom:
        return capi.Geom(self.slices, self.rows, self.cols, self.channels, float(self.pxl_dx), float(self.pxl_dy),
                         float(self.pxl_dz))


def _as_f32(vol: demons_volume) -> np.ndarray:
    return np.ascontiguousarray(vol.data, dtype=np.float32)


class DemonsEngine:
    """Device-resident engine: the replacement of the reference's `sycl_demons_engine` (src/Alignment_Demons.cc:52-704).

    Method names follow the reference engine's methods so stage-level parity tests read like the reference."""

    def __init__(self, fixed: demons_volume, moving: demons_volume, params: AlignViaDemonsParams, stream: int = 0):
        if fixed.data.shape != moving.data.shape:
            raise ValueError("fixed and moving volumes must have the same shape (resample first)")
        self._create(fixed.data.shape[:3], (fixed.pxl_dx, fixed.pxl_dy, fixed.pxl_dz), fixed.channels, params, stream)
        self.load(_as_f32(fixed), _as_f32(moving))

    @classmethod
    def empty(cls, shape, spacing, params: AlignViaDemonsParams, channels: int = 1, stream: int = 0):
        """Engine with allocated but unfilled buffers; fill with load()/load_device().  `stream` is a cudaStream_t
        handle (e.g. torch.cuda.current_stream().cuda_stream); 0 lets the library create its own stream."""
        self = cls.__new__(cls)
        self._create(tuple(int(v) for v in shape), spacing, channels, params, stream)
        return self

    def _create(self, shape, spacing, channels, params, stream):
        self._lib = capi.lib()
        self.params = params
        self.shape = tuple(shape)
        self.channels = int(channels)
        self.n_voxels = int(np.prod(self.shape))
        self._geom = capi.Geom(self.shape[0], self.shape[1], self.shape[2], self.channels, float(spacing[0]),
                               float(spacing[1]), float(spacing[2]))
        self._cparams = params.to_c()
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        capi.check(self._lib.dcma_b200_engine_create(C.byref(self._h), C.byref(self._geom), C.byref(self._cparams),
                                                     C.c_void_p(stream), err, len(err)), err)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dcma_b200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- data movement ------------------------------------------------------------------------------------------
    def load(self, fixed: np.ndarray, moving: np.ndarray):
        """Host float32 arrays -> HBM (reference: allocate_device_memory's copies, cc:181-186)."""
        fixed = np.ascontiguousarray(fixed, dtype=np.float32)
        moving = np.ascontiguousarray(moving, dtype=np.float32)
        assert fixed.size == self.n_voxels * self.channels == moving.size
        capi.check(self._lib.dcma_b200_engine_load(self._h, fixed.ctypes.data, moving.ctypes.data, 0))

    def load_device(self, fixed_ptr: int, moving_ptr: int):
        """Device pointers (e.g. torch tensors' data_ptr()) of float32 volumes on the engine's device."""
        capi.check(self._lib.dcma_b200_engine_load(self._h, C.c_void_p(fixed_ptr), C.c_void_p(moving_ptr), 1))

    def get(self, name: str) -> np.ndarray:
        which = {"fixed": 0, "moving": 1, "warped": 2, "gradient": 3, "update": 4, "deformation": 5}[name]
        if which <= 2:
            out = np.empty(self.shape + (self.channels,), dtype=np.float32)
        else:
            out = np.empty(self.shape + (3,), dtype=np.float64)
        capi.check(self._lib.dcma_b200_engine_get(self._h, which, out.ctypes.data))
        return out

    def set(self, name: str, arr: np.ndarray):
        which = {"fixed": 0, "moving": 1, "warped": 2, "update": 4, "deformation": 5}[name]
        if which <= 2:
            arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(self.shape + (self.channels,))
        else:
            arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(self.shape + (3,))
        capi.check(self._lib.dcma_b200_engine_set(self._h, which, arr.ctypes.data))

    def export_deformation_device(self, device_ptr: int):
        capi.check(self._lib.dcma_b200_engine_export_deformation_device(self._h, C.c_void_p(device_ptr)))

    def sync(self):
        capi.check(self._lib.dcma_b200_engine_sync(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.dcma_b200_engine_device_bytes(self._h))

    # -- the loop -----------------------------------------------------------------------------------------------
    def run(self, max_iterations: Optional[int] = None, profile: bool = False):
        """AlignViaDemons' iteration loop (cc:980-995). Returns (mse_history, stats)."""
        n = int(self.params.max_iterations if max_iterations is None else max_iterations)
        hist = np.full(max(n, 1), np.nan, dtype=np.float64)
        stats = capi.RunStats()
        capi.check(self._lib.dcma_b200_engine_run(self._h, n, 1 if profile else 0, hist.ctypes.data, C.byref(stats)))
        return hist[:stats.iterations_done].copy(), stats

    # -- reference engine methods, one at a time ------------------------------------------------------------------
    def _stage(self, stage, sigma=0.0):
        mse, n = C.c_double(0.0), C.c_int64(0)
        capi.check(self._lib.dcma_b200_engine_stage(self._h, stage, float(sigma), C.byref(mse), C.byref(n)))
        return mse.value, n.value

    def compute_update_and_mse(self):
        return self._stage(capi.STAGE_FORCE)

    def smooth_update(self, sigma_mm):
        self._stage(capi.STAGE_SMOOTH_U, sigma_mm)

    def smooth_deformation(self, sigma_mm):
        self._stage(capi.STAGE_SMOOTH_D, sigma_mm)

    def add_update(self):
        self._stage(capi.STAGE_ADD)

    def add_update_diffeomorphic(self):
        self._stage(capi.STAGE_COMPOSE)

    def warp_moving(self):
        self._stage(capi.STAGE_WARP)


def histogram_match(source: np.ndarray, reference: np.ndarray, num_bins: int = 256, outlier_fraction: float = 0.01,
                    device: int = -1):
    """AlignViaDemonsHelpers::histogram_match (reference src/Alignment_Demons.cc:760-891) on the GPU, bit-identical.
    Returns (matched array shaped like `source`, mapped flag)."""
    lib = capi.lib()
    src = np.ascontiguousarray(source, dtype=np.float32)
    ref = np.ascontiguousarray(reference, dtype=np.float32)
    if src.size == 0 or ref.size == 0:
        raise ValueError("Cannot perform histogram matching: image collection is empty")
    out = np.empty_like(src)
    mapped = C.c_int32(0)
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_histogram_match(src.ctypes.data, src.size, ref.ctypes.data, ref.size, int(num_bins),
                                             float(outlier_fraction), out.ctypes.data, C.byref(mapped), device, err,
                                             len(err)), err)
    return out, bool(mapped.value)


def make_grid(shape, spacing, origin=(0.0, 0.0, 0.0), channels=1, ex=(1.0, 0.0, 0.0), ey=(0.0, 1.0, 0.0),
              ez=(0.0, 0.0, 1.0)) -> capi.Grid:
    """A regular grid in world space: voxel (z, y, x) at origin + x*dx*ex + y*dy*ey + z*dz*ez; spacing = (dx, dy, dz)."""
    g = capi.Grid()
    g.geom = capi.Geom(int(shape[0]), int(shape[1]), int(shape[2]), int(channels), float(spacing[0]), float(spacing[1]),
                       float(spacing[2]))
    for i in range(3):
        g.origin[i], g.ex[i], g.ey[i], g.ez[i] = float(origin[i]), float(ex[i]), float(ey[i]), float(ez[i])
    return g


def warp_image_on_grid(field_grid: capi.Grid, deformation: np.ndarray, image_grid: capi.Grid, image: np.ndarray,
                       mask: Optional[np.ndarray] = None, oob_value: float = 0.0, device: int = -1) -> np.ndarray:
    """WarpImages' voxel body for a deformation_field transform when the image (e.g. an RT dose grid) lives on another
    regular grid than the field (reference src/Operations/WarpImages.cc:418-456)."""
    lib = capi.lib()
    D = np.ascontiguousarray(deformation, dtype=np.float64)
    img = np.ascontiguousarray(image, dtype=np.float32)
    out = np.empty_like(img)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_warp_image_grid(C.byref(field_grid), D.ctypes.data, C.byref(image_grid), img.ctypes.data,
                                             None if m is None else m.ctypes.data, float(oob_value), out.ctypes.data,
                                             device, err, len(err)), err)
    return out


def warp_to_grid(field_grid: capi.Grid, deformation: np.ndarray, source_grid: capi.Grid, source: np.ndarray, out_grid: capi.Grid,
                 placeholder: np.ndarray, mask: Optional[np.ndarray] = None, channel: int = -1, oob_value: float = 0.0,
                 device: int = -1) -> np.ndarray:
    """WarpImages' voxel body with the operation's three stacks (reference src/Operations/WarpImages.cc:385-456): the voxels
    of `placeholder` (the copy of the reference image array, on `out_grid`) that `mask` selects take
    source(p + D(p)); the others, and channels other than `channel` (>= 0), are returned unchanged."""
    lib = capi.lib()
    D = np.ascontiguousarray(deformation, dtype=np.float64)
    src = np.ascontiguousarray(source, dtype=np.float32)
    out = np.array(placeholder, dtype=np.float32, order="C", copy=True)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_warp_to_grid(C.byref(field_grid), D.ctypes.data, C.byref(source_grid), src.ctypes.data,
                                          C.byref(out_grid), None if m is None else m.ctypes.data, int(channel), float(oob_value),
                                          out.ctypes.data, device, err, len(err)), err)
    return out


CHUNK_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64)


def register_stacks(params: "AlignViaDemonsParams", fixed_grid: capi.Grid, fixed: np.ndarray, moving_grid: capi.Grid,
                    moving: np.ndarray, on_chunk=None):
    """dcma_b200_demons_register_stacks: the moving stack (on its own regular grid) is resampled onto the fixed grid on the
    device, registered, and the field comes back in chunks; `on_chunk(voxel_begin, voxel_end)` is called per landed chunk.
    Returns (deformation (S,R,C,3) float64, mse_history, stats)."""
    lib = capi.lib()
    f = np.ascontiguousarray(fixed, dtype=np.float32)
    m = np.ascontiguousarray(moving, dtype=np.float32)
    g = fixed_grid.geom
    out = np.empty((g.slices, g.rows, g.cols, 3), dtype=np.float64)
    cp = params.to_c()
    n_hist = 4 * int(params.max_iterations) + sum(int(v) for v in params.pyramid_iterations)
    hist = np.zeros(max(n_hist, 1), dtype=np.float64)
    st = capi.RunStats()
    err = C.create_string_buffer(512)
    cb = CHUNK_FN(lambda user, a, b: on_chunk(int(a), int(b))) if on_chunk is not None else None
    capi.check(lib.dcma_b200_demons_register_stacks(C.byref(fixed_grid), f.ctypes.data, C.byref(moving_grid), m.ctypes.data,
                                                    C.byref(cp), out.ctypes.data, hist.ctypes.data, C.byref(st),
                                                    C.cast(cb, C.c_void_p) if cb is not None else None, None, err, len(err)), err)
    return out, hist[:st.iterations_done], st


VFIELD_KINDS = {"sinusoidal_v1": 0, "vortex_v1": 1}


def virtual_field_v1(kind: str, shape=(100, 512, 512), bounds=(0.0, 512.0, 0.0, 512.0, 0.0, 100.0), device: int = -1):
    """The synthetic deformation field of GenerateVirtualDataSinusoidalFieldV1 / GenerateVirtualDataVortexFieldV1 (reference
    src/Operations/GenerateVirtualData{Sinusoidal,Vortex}FieldV1.cc) on the operations' grid: `shape` voxels across `bounds`,
    in-plane positions at pxl * index (img.position of a stack anchored at the origin), slices at z_min + pxl_dz * (z + 1/2).
    Returns (field (S,R,C,3) float64 mm, grid, (min, max) magnitude before normalisation)."""
    lib = capi.lib()
    S, R, Cn = (int(v) for v in shape)
    b = np.asarray(bounds, dtype=np.float64)
    sp = ((b[1] - b[0]) / Cn, (b[3] - b[2]) / R, (b[5] - b[4]) / S)
    grid = make_grid((S, R, Cn), sp, origin=(0.0, 0.0, b[4] + 0.5 * sp[2]), channels=3)
    out = np.empty((S, R, Cn, 3), dtype=np.float64)
    ext = np.zeros(2, dtype=np.float64)
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_virtual_field(VFIELD_KINDS[kind], C.byref(grid), b.ctypes.data, out.ctypes.data, ext.ctypes.data,
                                           device, err, len(err)), err)
    return out, grid, (float(ext[0]), float(ext[1]))


def resample_to_grid(src_grid: capi.Grid, image: np.ndarray, dst_grid: capi.Grid, oob_value: float = float("nan"),
                     device: int = -1) -> np.ndarray:
    """resample_image_to_reference_grid (reference src/Alignment_Demons.cc:709-755) for regular grids."""
    lib = capi.lib()
    img = np.ascontiguousarray(image, dtype=np.float32)
    ch = int(src_grid.geom.channels)
    out = np.empty((dst_grid.geom.slices, dst_grid.geom.rows, dst_grid.geom.cols) + ((ch,) if img.ndim == 4 else ()), dtype=np.float32)
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_resample_to_grid(C.byref(src_grid), img.ctypes.data, C.byref(dst_grid), float(oob_value),
                                              out.ctypes.data, device, err, len(err)), err)
    return out


def restrict_image(image: demons_volume, device: int = -1) -> demons_volume:
    """Next-coarser pyramid level of an image (extension a12; definition in csrc/pyramid.cu)."""
    lib = capi.lib()
    img = _as_f32(image)
    g, gc = image.geom(), capi.Geom()
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_restrict_image(C.byref(g), None, C.byref(gc), None, device, err, len(err)), err)
    out = np.empty((gc.slices, gc.rows, gc.cols, gc.channels), dtype=np.float32)
    capi.check(lib.dcma_b200_restrict_image(C.byref(g), img.ctypes.data, C.byref(gc), out.ctypes.data, device, err, len(err)), err)
    return demons_volume(out, gc.pxl_dx, gc.pxl_dy, gc.pxl_dz)


def prolong_field(coarse_field: np.ndarray, fine: demons_volume, device: int = -1) -> np.ndarray:
    """Trilinear prolongation of a displacement field from the next-coarser level onto the grid of `fine`."""
    lib = capi.lib()
    cf = np.ascontiguousarray(coarse_field, dtype=np.float64)
    out = np.empty(fine.data.shape[:3] + (3,), dtype=np.float64)
    g = fine.geom()
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_prolong_field(C.byref(g), cf.ctypes.data, out.ctypes.data, device, err, len(err)), err)
    return out


def slab_partition(slices: int, world: int, rank: int, halo_planes: int):
    """(zoff, local, own0, own1) of one z-slab: pure host arithmetic of the library (no GPU needed)."""
    v = [C.c_int64(0) for _ in range(4)]
    capi.check(capi.lib().dcma_b200_slab_partition(int(slices), int(world), int(rank), int(halo_planes),
                                                   *[C.byref(x) for x in v]))
    return tuple(int(x.value) for x in v)


class DemonsSlabGroup:
    """One registration sharded into z-slabs (BASELINE config 4): `DemonsSlabGroup.local(...)` drives
    several engines from this process (peer copies), `DemonsSlabGroup.nccl(...)` is one rank of a multi-process run
    (one process per GPU; halos by NCCL).  The reference has no counterpart (one address space, cc:171-187)."""

    def __init__(self):
        raise TypeError("use DemonsSlabGroup.local(...) or DemonsSlabGroup.nccl(...)")

    @classmethod
    def _new(cls, shape, spacing, params, channels):
        self = cls.__new__(cls)
        self._lib = capi.lib()
        self.params = params
        self.shape = tuple(int(v) for v in shape)
        self.channels = int(channels)
        self._geom = capi.Geom(self.shape[0], self.shape[1], self.shape[2], self.channels, float(spacing[0]),
                               float(spacing[1]), float(spacing[2]))
        self._cparams = params.to_c()
        self._h = C.c_void_p()
        return self

    @classmethod
    def local(cls, shape, spacing, params: AlignViaDemonsParams, devices, halo_planes: int = -1, channels: int = 1):
        self = cls._new(shape, spacing, params, channels)
        dev = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        err = C.create_string_buffer(512)
        capi.check(self._lib.dcma_b200_slab_group_create_local(C.byref(self._h), C.byref(self._geom), C.byref(self._cparams),
                                                               dev, len(devices), int(halo_planes), err, len(err)), err)
        return self

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * 128)()
        err = C.create_string_buffer(512)
        capi.check(capi.lib().dcma_b200_nccl_unique_id(buf, err, len(err)), err)
        return bytes(buf)

    @classmethod
    def nccl(cls, shape, spacing, params: AlignViaDemonsParams, unique_id: bytes, rank: int, world: int, device: int,
             halo_planes: int = -1, channels: int = 1):
        self = cls._new(shape, spacing, params, channels)
        buf = (C.c_ubyte * 128)(*unique_id)
        err = C.create_string_buffer(512)
        capi.check(self._lib.dcma_b200_slab_group_create_nccl(C.byref(self._h), C.byref(self._geom), C.byref(self._cparams),
                                                              buf, int(rank), int(world), int(device), int(halo_planes),
                                                              err, len(err)), err)
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dcma_b200_slab_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, fixed: np.ndarray, moving: np.ndarray):
        """WHOLE volumes (host float32)."""
        fixed = np.ascontiguousarray(fixed, dtype=np.float32)
        moving = np.ascontiguousarray(moving, dtype=np.float32)
        capi.check(self._lib.dcma_b200_slab_group_load(self._h, fixed.ctypes.data, moving.ctypes.data, 0))

    def load_device(self, fixed_ptr: int, moving_ptr: int):
        capi.check(self._lib.dcma_b200_slab_group_load(self._h, C.c_void_p(fixed_ptr), C.c_void_p(moving_ptr), 1))

    def run(self, max_iterations: Optional[int] = None):
        n = int(self.params.max_iterations if max_iterations is None else max_iterations)
        hist = np.full(max(n, 1), np.nan, dtype=np.float64)
        stats = capi.RunStats()
        capi.check(self._lib.dcma_b200_slab_group_run(self._h, n, hist.ctypes.data, C.byref(stats)))
        return hist[:stats.iterations_done].copy(), stats

    def owned_range(self):
        z0, z1 = C.c_int64(0), C.c_int64(0)
        capi.check(self._lib.dcma_b200_slab_group_range(self._h, C.byref(z0), C.byref(z1)))
        return int(z0.value), int(z1.value)

    def get_deformation(self) -> np.ndarray:
        """Deformation of the slices this process holds: (z1 - z0, rows, cols, 3) float64 mm."""
        z0, z1 = self.owned_range()
        out = np.empty((z1 - z0, self.shape[1], self.shape[2], 3), dtype=np.float64)
        capi.check(self._lib.dcma_b200_slab_group_get_deformation(self._h, out.ctypes.data))
        return out

    @property
    def halo_bytes(self) -> int:
        return int(self._lib.dcma_b200_slab_group_halo_bytes(self._h))

    TRANSPORTS = {0: "in-process peer copies + events", 1: "NCCL send/recv", 2: "IPC pull (copy engines) + NCCL tokens",
                  3: "IPC push (copy engines) + flag counters (stream memory operations)"}

    def exchange_stats(self) -> dict:
        """Halo-exchange statistics of this process's last run (see dcma_b200_slab_group_exchange_stats)."""
        ms, n, ov, tr, hp = C.c_double(0), C.c_int64(0), C.c_int32(0), C.c_int32(0), C.c_int64(0)
        capi.check(self._lib.dcma_b200_slab_group_exchange_stats(self._h, C.byref(ms), C.byref(n), C.byref(ov), C.byref(tr),
                                                                 C.byref(hp)))
        return {"exchange_ms": float(ms.value), "exchanges": int(n.value), "overlapped": bool(ov.value),
                "transport": self.TRANSPORTS.get(int(tr.value), str(tr.value)), "halo_planes": int(hp.value)}


def AlignViaDemons(params: AlignViaDemonsParams, moving: demons_volume, stationary: demons_volume,
                   return_stats: bool = False):
    """Drop-in for the reference's AlignViaDemons (src/Alignment_Demons.h:233-236) at the `demons_volume` level.

    `moving` must already be on the stationary grid (the reference's step 1, cc:964; identity when grids coincide).
    Returns the deformation volume (slices, rows, cols, 3) float64 in mm, or None on failure -- the reference returns
    std::nullopt after logging a warning (cc:953-956, 1001-1004)."""
    if moving.data.size == 0 or stationary.data.size == 0:
        print("Unable to perform demons alignment: an image array is empty", file=sys.stderr)
        return (None, None, None) if return_stats else None
    try:
        lib = capi.lib()
        if moving.data.shape != stationary.data.shape:
            raise ValueError("moving image is not on the stationary grid")
        f, m = _as_f32(stationary), _as_f32(moving)
        geom = stationary.geom()
        cp = params.to_c()
        D = np.empty(stationary.data.shape[:3] + (3,), dtype=np.float64)
        # one entry per iteration of every pyramid level (coarse levels first)
        n_hist = int(params.max_iterations)
        for l in range(1, int(params.pyramid_levels)):
            extra = int(params.pyramid_iterations[l - 1]) if l - 1 < len(params.pyramid_iterations) else 0
            n_hist += extra if extra > 0 else int(params.max_iterations)
        hist = np.full(max(n_hist, 1), np.nan, dtype=np.float64)
        stats = capi.RunStats()
        err = C.create_string_buffer(512)
        capi.check(lib.dcma_b200_demons_register(C.byref(geom), f.ctypes.data, m.ctypes.data, C.byref(cp),
                                                 D.ctypes.data, hist.ctypes.data, C.byref(stats), err, len(err)), err)
        out = demons_volume(D, stationary.pxl_dx, stationary.pxl_dy, stationary.pxl_dz)
        if return_stats:
            return out, hist[:stats.iterations_done].copy(), stats
        return out
    except (capi.DcmaError, ValueError) as e:
        print(f"Demons registration failed: {e}", file=sys.stderr)
    return (None, None, None) if return_stats else None


def warp_image_with_field(image: demons_volume, deformation: np.ndarray, oob_value: Optional[float] = None,
                          device: int = -1) -> np.ndarray:
    """warp_moving semantics (cc:467-550) for an image on the field's grid: out(p) = image(p + D(p)).

    oob_value=None keeps NaN where the reference writes NaN; WarpImages uses 0.0 (src/Operations/WarpImages.cc:197)."""
    lib = capi.lib()
    img = _as_f32(image)
    D = np.ascontiguousarray(deformation, dtype=np.float64).reshape(image.data.shape[:3] + (3,))
    out = np.empty_like(img)
    geom = image.geom()
    err = C.create_string_buffer(512)
    capi.check(lib.dcma_b200_warp_image(C.byref(geom), img.ctypes.data, D.ctypes.data,
                                        0.0 if oob_value is None else float(oob_value),
                                        0 if oob_value is None else 1, out.ctypes.data, device, err, len(err)), err)
    return out
