"""Deterministic synthetic CT-like phantom pair with a known smooth sinusoidal warp.

The pair is generated analytically on the fixed grid, so the reference's pre-step
`resample_image_to_reference_grid` (src/Alignment_Demons.cc:709-755) is the identity and no NaNs exist at t=0.
The warp is modelled on the reference's own generator (src/Operations/GenerateVirtualDataSinusoidalFieldV1.cc:77-134).

Fixed  F(p)  = phantom(p)            HU: -1000 air, soft-tissue ellipsoid with sigmoid edge, +300 / -700 HU Gaussian
                                      blobs, +900 HU rod, and a 20 HU sinusoidal texture EVERYWHERE (so |grad F| is
                                      never tiny-but-nonzero over large regions).
Moving M(p)  = phantom(p + w(p)/pxl)  with w = A_w * ( sin(2pi y/ly) cos(2pi z/lz),
                                                       sin(2pi x/lx) cos(2pi z/lz),
                                                       sin(2pi x/lx) sin(2pi y/ly) ),  l = extent / (2, 3, 1.5).
Pull convention warped(p) = moving(p + D(p)): the ideal recovered field satisfies D(p) = -w(p + D(p)) ~ -w(p).

Pure torch so the same code fills 128^3 on the CPU and 1024^3 directly in HBM.
"""
from __future__ import annotations

import math

import torch


# Fine structure of the phantom.  "periodic" (the benchmark phantom): three plane waves of ~7-9 voxels wavelength, 20 HU
# each -- every voxel gets a well-conditioned force, but the only information below the organ scale is
# periodic, so displacements beyond half a wavelength (~3.5 voxels) are ambiguous.  "multiscale": the same three wave
# directions at four octaves (wavelengths ~7 ... ~70 voxels, amplitude growing with wavelength, incommensurate phases) --
# structure at every scale, as anatomy has, which is what a multi-resolution registration feeds on (BASELINE config 3).
TEXTURES = ("periodic", "multiscale")


def _texture(x, y, z, ph, texture):
    if texture == "periodic":
        return 20.0 * (torch.sin(0.9 * x + 0.3 * y + ph) + torch.sin(0.7 * y + 0.4 * z) + torch.sin(0.8 * z + 0.5 * x))
    if texture != "multiscale":
        raise ValueError(f"texture must be one of {TEXTURES}")
    out = 0.0
    for k, amp in enumerate((20.0, 30.0, 45.0, 70.0)):
        s = 0.5 ** k * (1.0 + 0.07 * k)  # not exact octaves: no common period
        out = out + amp * (torch.sin(s * (0.9 * x + 0.3 * y) + ph + 1.3 * k) + torch.sin(s * (0.7 * y + 0.4 * z) + 2.1 * k) +
                           torch.sin(s * (0.8 * z + 0.5 * x) + 0.7 * k))
    return out


def _phantom_hu(x, y, z, shape, case: int = 0, texture: str = "periodic"):
    """x, y, z: float64 tensors of voxel-index coordinates (x=col, y=row, z=slice). shape = (slices, rows, cols)."""
    nz, ny, nx = (float(v) for v in shape)
    cx, cy, cz = (nx - 1.0) / 2.0, (ny - 1.0) / 2.0, (nz - 1.0) / 2.0
    ph = 0.37 * float(case)  # per-case variation (batch mode, BASELINE config 5)
    small = min(nx, ny, nz)

    e = torch.sqrt(((x - cx) / (0.40 * nx)) ** 2 + ((y - cy) / (0.32 * ny)) ** 2 + ((z - cz) / (0.45 * nz)) ** 2)
    body = torch.sigmoid((1.0 - e) / 0.03)
    hu = -1000.0 + 1000.0 * body

    def blob(amp, fx, fy, fz, s):
        bx, by, bz = cx + fx * nx, cy + fy * ny, cz + fz * nz
        r2 = (x - bx) ** 2 + (y - by) ** 2 + (z - bz) ** 2
        return amp * torch.exp(-0.5 * r2 / (s * small) ** 2)

    hu = hu + body * blob(300.0, 0.15 * math.cos(ph), -0.05 + 0.10 * math.sin(ph), 0.10, 0.08)
    hu = hu + body * blob(-700.0, -0.15 * math.cos(ph), 0.08, -0.12 + 0.05 * math.sin(ph), 0.07)
    rho2 = (x - cx) ** 2 + (y - (cy + 0.18 * ny)) ** 2
    hu = hu + body * 900.0 * torch.exp(-0.5 * rho2 / (0.03 * small) ** 2)
    hu = hu + _texture(x, y, z, ph, texture)
    return hu


def known_warp_mm(x, y, z, shape, amplitude_mm: float = 2.0, case: int = 0):
    """The known warp w(p) in mm at voxel-index coordinates (x, y, z). Returns (wx, wy, wz)."""
    nz, ny, nx = (float(v) for v in shape)
    lx, ly, lz = nx / 2.0, ny / 3.0, max(nz / 1.5, 1.0)
    ph = 0.37 * float(case)
    two_pi = 2.0 * math.pi
    sx, sy = torch.sin(two_pi * x / lx + ph), torch.sin(two_pi * y / ly)
    czs = torch.cos(two_pi * z / lz)
    return amplitude_mm * sy * czs, amplitude_mm * sx * czs, amplitude_mm * sx * sy


def make_pair(shape, spacing=(1.0, 1.0, 1.0), amplitude_mm: float = 2.0, case: int = 0, device="cpu",
              z_range=None, chunk: int = 16, texture: str = "periodic"):
    """Returns (fixed, moving) float32 tensors of shape (slices, rows, cols) [or the z_range sub-slab of it].

    spacing = (pxl_dx, pxl_dy, pxl_dz) mm. For a single-slice volume the z component of the warp is dropped."""
    nz, ny, nx = (int(v) for v in shape)
    z0, z1 = (0, nz) if z_range is None else (int(z_range[0]), int(z_range[1]))
    dev = torch.device(device)
    fixed = torch.empty((z1 - z0, ny, nx), dtype=torch.float32, device=dev)
    moving = torch.empty_like(fixed)
    ys = torch.arange(ny, dtype=torch.float64, device=dev).view(1, ny, 1)
    xs = torch.arange(nx, dtype=torch.float64, device=dev).view(1, 1, nx)
    for a in range(z0, z1, chunk):
        b = min(a + chunk, z1)
        zs = torch.arange(a, b, dtype=torch.float64, device=dev).view(b - a, 1, 1)
        x, y, z = torch.broadcast_tensors(xs, ys, zs)
        fixed[a - z0:b - z0] = _phantom_hu(x, y, z, shape, case, texture).to(torch.float32)
        wx, wy, wz = known_warp_mm(x, y, z, shape, amplitude_mm, case)
        if nz == 1:
            wz = torch.zeros_like(wz)
        moving[a - z0:b - z0] = _phantom_hu(x + wx / spacing[0], y + wy / spacing[1], z + wz / spacing[2],
                                            shape, case, texture).to(torch.float32)
    return fixed, moving


def make_dose_grid(shape, device="cpu", chunk: int = 16):
    """Smooth 0-70 Gy blob on the same grid (BASELINE config 3: 'warp of a co-registered RT dose grid')."""
    nz, ny, nx = (int(v) for v in shape)
    dev = torch.device(device)
    out = torch.empty((nz, ny, nx), dtype=torch.float32, device=dev)
    ys = torch.arange(ny, dtype=torch.float64, device=dev).view(1, ny, 1)
    xs = torch.arange(nx, dtype=torch.float64, device=dev).view(1, 1, nx)
    for a in range(0, nz, chunk):
        b = min(a + chunk, nz)
        zs = torch.arange(a, b, dtype=torch.float64, device=dev).view(b - a, 1, 1)
        r2 = ((xs - 0.55 * nx) / (0.18 * nx)) ** 2 + ((ys - 0.45 * ny) / (0.15 * ny)) ** 2 \
            + ((zs - 0.5 * nz) / (0.2 * nz)) ** 2
        out[a:b] = (70.0 * torch.exp(-0.5 * r2)).to(torch.float32)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Bit-reproducible variant.  `make_pair` calls sin/exp/sigmoid, whose last bit differs between libm, SLEEF and CUDA, so a
# pair made on one machine cannot be hashed against one made on another.  `make_pair_exact` builds the same kind of
# phantom (air, sigmoid-edged ellipsoid, two blobs, a rod, texture everywhere, the same family of sinusoidal warp) from
# IEEE-exact float64 operations only -- + - * / sqrt floor, one torch kernel per operation, so nothing is contracted into
# an FMA -- and is therefore bit-identical on any CPU and on the GPU.  Used by the full-size parity fixtures
# (tests/golden/make_full_size_fixtures.py): the committed sha256 of the reference's result is only meaningful if the
# GPU box regenerates exactly the same inputs.
# ----------------------------------------------------------------------------------------------------------------------
_TWO_PI = 6.283185307179586
_INV_TWO_PI = 0.15915494309189535


def _sin_exact(t):
    """A sine made of exact operations: reduce to [-pi, pi), then an odd degree-11 polynomial (|error| < 5e-4 at the
    interval ends, far below what the phantom needs; smooth and periodic is all that matters)."""
    k = torch.floor(t * _INV_TWO_PI + 0.5)
    r = t - k * _TWO_PI
    r2 = r * r
    p = r2 * (-2.505210838544172e-08) + 2.755731922398589e-06
    p = p * r2 + (-1.984126984126984e-04)
    p = p * r2 + 8.333333333333333e-03
    p = p * r2 + (-1.666666666666667e-01)
    p = p * r2 + 1.0
    return r * p


def _cos_exact(t):
    return _sin_exact(t + 1.5707963267948966)


def _bump_exact(q):
    """Stand-in for exp(-q/2): (1 + q/8)^-4, same value/slope at 0, positive, smooth, algebraic tail."""
    b = 1.0 / (q * 0.125 + 1.0)
    b = b * b
    return b * b


def _phantom_hu_exact(x, y, z, shape):
    nz, ny, nx = (float(v) for v in shape)
    cx, cy, cz = (nx - 1.0) / 2.0, (ny - 1.0) / 2.0, (nz - 1.0) / 2.0
    small = min(nx, ny, nz)
    # NB: never `tensor / python_scalar` here -- torch's CUDA kernel multiplies by the rounded reciprocal while the CPU
    # kernel divides, a 1-ulp difference that flips a float32 rounding once in ~1e8 voxels.  Multiply by a reciprocal
    # computed once in Python instead (the same double on every device).
    ex, ey, ez = (x - cx) * (1.0 / (0.40 * nx)), (y - cy) * (1.0 / (0.32 * ny)), (z - cz) * (1.0 / (0.45 * nz))
    e = torch.sqrt(ex * ex + ey * ey + ez * ez)
    u = (1.0 - e) * (1.0 / 0.06)
    body = (u / torch.sqrt(u * u + 1.0)) * 0.5 + 0.5  # algebraic sigmoid
    hu = body * 1000.0 - 1000.0

    def blob(amp, fx, fy, fz, s):
        bx, by, bz = cx + fx * nx, cy + fy * ny, cz + fz * nz
        dx, dy, dz = x - bx, y - by, z - bz
        return _bump_exact((dx * dx + dy * dy + dz * dz) * (1.0 / ((s * small) ** 2))) * amp

    hu = hu + body * blob(300.0, 0.15, -0.05, 0.10, 0.08)
    hu = hu + body * blob(-700.0, -0.15, 0.08, -0.12, 0.07)
    rx, ry = x - cx, y - (cy + 0.18 * ny)
    hu = hu + body * (_bump_exact((rx * rx + ry * ry) * (1.0 / ((0.03 * small) ** 2))) * 900.0)
    tex = _sin_exact(x * 0.9 + y * 0.3) + _sin_exact(y * 0.7 + z * 0.4) + _sin_exact(z * 0.8 + x * 0.5)
    return hu + tex * 20.0


def known_warp_mm_exact(x, y, z, shape, amplitude_mm: float = 2.0):
    nz, ny, nx = (float(v) for v in shape)
    lx, ly, lz = nx / 2.0, ny / 3.0, max(nz / 1.5, 1.0)
    sx, sy = _sin_exact(x * (_TWO_PI / lx)), _sin_exact(y * (_TWO_PI / ly))
    czs = _cos_exact(z * (_TWO_PI / lz))
    return (sy * czs) * amplitude_mm, (sx * czs) * amplitude_mm, (sx * sy) * amplitude_mm


def make_pair_exact(shape, spacing=(1.0, 1.0, 1.0), amplitude_mm: float = 2.0, device="cpu", chunk: int = 16):
    """Bit-reproducible (fixed, moving) float32 pair: identical bytes on every CPU and on the GPU (see above)."""
    nz, ny, nx = (int(v) for v in shape)
    dev = torch.device(device)
    fixed = torch.empty((nz, ny, nx), dtype=torch.float32, device=dev)
    moving = torch.empty_like(fixed)
    ys = torch.arange(ny, dtype=torch.float64, device=dev).view(1, ny, 1)
    xs = torch.arange(nx, dtype=torch.float64, device=dev).view(1, 1, nx)
    for a in range(0, nz, chunk):
        b = min(a + chunk, nz)
        zs = torch.arange(a, b, dtype=torch.float64, device=dev).view(b - a, 1, 1)
        x, y, z = (t.contiguous() for t in torch.broadcast_tensors(xs, ys, zs))
        fixed[a:b] = _phantom_hu_exact(x, y, z, shape).to(torch.float32)
        wx, wy, wz = known_warp_mm_exact(x, y, z, shape, amplitude_mm)
        if nz == 1:
            wz = wz * 0.0
        moving[a:b] = _phantom_hu_exact(x + wx * (1.0 / spacing[0]), y + wy * (1.0 / spacing[1]),
                                        z + wz * (1.0 / spacing[2]), shape).to(torch.float32)
    return fixed, moving
