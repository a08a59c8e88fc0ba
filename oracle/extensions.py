"""TEST INFRASTRUCTURE ONLY: CPU restatement (numpy, fp64) of the two EXTENSIONS that have no counterpart in the
reference -- the symmetric Demons force and the multi-resolution pyramid (row a12).

**Parity unpinned**: `AlignViaDemonsParams` (/root/reference/src/Alignment_Demons.h:20-61) has neither option and the
reference's tests hold nothing for them, so there is no reference output to compare with.  What IS pinned: every stage
these loops are assembled from (gradient rule, smoothing, composition, warp) is the pinned oracle engine of
`oracle/__init__.py`; only the few lines below (mean of two gradients; block mean; trilinear prolongation) are ours, and
they restate the definitions documented in dicomautomaton_b200/csrc/pyramid.cu and demons_kernels.cuh.
"""
from __future__ import annotations

import numpy as np

from . import OracleEngine, make_params


def gradient_rule(img: np.ndarray) -> np.ndarray:
    """compute_gradient's rule (reference src/Alignment_Demons.cc:232-265) applied to a (S,R,C) float image:
    (v[+1] - v[-1]) / (index distance) with clamped neighbours, 0 where a neighbour is non-finite or the extent is 1."""
    v = img.astype(np.float64)
    out = np.zeros(v.shape + (3,), dtype=np.float64)
    for comp, axis in ((0, 2), (1, 1), (2, 0)):  # x <-> columns, y <-> rows, z <-> slices
        n = v.shape[axis]
        if n <= 1:
            continue
        idx = np.arange(n)
        lo, hi = np.maximum(idx - 1, 0), np.minimum(idx + 1, n - 1)
        a, b = np.take(v, hi, axis=axis), np.take(v, lo, axis=axis)
        shape = [1, 1, 1]
        shape[axis] = n
        dist = (hi - lo).astype(np.float64).reshape(shape)
        with np.errstate(invalid="ignore"):
            g = (a - b) / dist
        g[~(np.isfinite(a) & np.isfinite(b))] = 0.0
        out[..., comp] = g
    return out


def symmetric_force(fixed, warped, grad_fixed, normalization_factor=1.0, max_update_magnitude=2.0):
    """compute_update_and_mse (cc:291-330) with J = (grad F + grad W) / 2 in place of grad F.  Returns (U, mse, n)."""
    f, w = fixed.astype(np.float64), warped.astype(np.float64)
    valid = np.isfinite(fixed) & np.isfinite(warped)
    J = (grad_fixed + gradient_rule(warped)) * 0.5
    with np.errstate(invalid="ignore", over="ignore"):
        diff = np.where(valid, f - w, 0.0)
        gmag = J[..., 0] * J[..., 0] + J[..., 1] * J[..., 1] + J[..., 2] * J[..., 2]
        denom = gmag + (diff * diff) / (normalization_factor + 1.0e-10)
        ok = valid & (denom > 1.0e-10)
        safe = np.where(ok, denom, 1.0)
        U = np.where(ok[..., None], diff[..., None] * J / safe[..., None], 0.0)
        mag = np.sqrt(U[..., 0] * U[..., 0] + U[..., 1] * U[..., 1] + U[..., 2] * U[..., 2])
        scale = np.where(mag > max_update_magnitude, max_update_magnitude / np.where(mag > 0, mag, 1.0), 1.0)
        U = np.where((mag > max_update_magnitude)[..., None], U * scale[..., None], U)
    n = int(valid.sum())
    mse = float((diff * diff)[valid].sum() / n) if n > 0 else 0.0
    return U, mse, n


def run_level(fixed, moving, spacing, iters, sigma_d, sigma_u, diffeo, symmetric, D0=None, normalization_factor=1.0,
              max_update_magnitude=2.0):
    """The loop of one level (threshold 0) assembled from the pinned oracle stages; D0 = warm start (mm)."""
    par = make_params(max_iterations=iters, convergence_threshold=0.0, deformation_field_smoothing_sigma=sigma_d,
                      update_field_smoothing_sigma=sigma_u, use_diffeomorphic=diffeo,
                      normalization_factor=normalization_factor, max_update_magnitude=max_update_magnitude)
    E = OracleEngine(fixed, moving, par, spacing, kind="port")
    hist = []
    try:
        if D0 is not None:
            E.set("deformation", D0)
            E.warp_moving()  # W := warp(M, D0); W := M only holds for D = 0
        G = E.get("gradient")
        for _ in range(iters):
            if symmetric:
                U, mse, _ = symmetric_force(E.get("fixed")[..., 0], E.get("warped")[..., 0], G, normalization_factor,
                                            max_update_magnitude)
                E.set("update", U)
            else:
                mse, _ = E.compute_update_and_mse()
            hist.append(mse)
            if diffeo and sigma_u > 0:
                E.smooth_update(sigma_u)
            if diffeo:
                E.add_update_diffeomorphic()
            else:
                E.add_update()
            if sigma_d > 0:
                E.smooth_deformation(sigma_d)
            E.warp_moving()
        return E.get("deformation"), np.array(hist)
    finally:
        E.close()


def coarsen(shape, spacing):
    """(coarse shape, coarse spacing (dx,dy,dz), halved flags (z,y,x)): axes with extent >= 8 are halved."""
    s, r, c = shape
    hz, hy, hx = s >= 8, r >= 8, c >= 8
    cs = ((s + 1) // 2 if hz else s, (r + 1) // 2 if hy else r, (c + 1) // 2 if hx else c)
    sp = (spacing[0] * (2.0 if hx else 1.0), spacing[1] * (2.0 if hy else 1.0), spacing[2] * (2.0 if hz else 1.0))
    return cs, sp, (hz, hy, hx)


_BLUR_W = np.array([0.011108996538242306, 0.1353352832366127, 0.6065306597126334, 1.0,
                    0.6065306597126334, 0.1353352832366127, 0.011108996538242306], dtype=np.float64)


def prefilter(img: np.ndarray):
    """Anti-aliasing low-pass before decimation (csrc/pyramid.cu): along every axis that is about to be halved, a 7-tap
    Gaussian (sigma = 1 voxel), weights renormalised over the in-grid finite taps, double accumulation in tap order,
    rounded to float after each axis; x, then y, then z."""
    _, _, (hz, hy, hx) = coarsen(img.shape, (1, 1, 1))
    out = img.astype(np.float32)
    for axis, halved in ((2, hx), (1, hy), (0, hz)):
        if not halved:
            continue
        n = out.shape[axis]
        src = out.astype(np.float64)
        fin = np.isfinite(src)
        vals = np.where(fin, src, 0.0)
        acc = np.zeros_like(src)
        wsum = np.zeros_like(src)
        for k in range(-3, 4):
            lo, hi = max(0, -k), min(n, n - k)  # output positions whose tap pos + k is inside the grid
            if lo >= hi:
                continue
            dst = [slice(None)] * 3
            frm = [slice(None)] * 3
            dst[axis] = slice(lo, hi)
            frm[axis] = slice(lo + k, hi + k)
            dst, frm = tuple(dst), tuple(frm)
            acc[dst] = acc[dst] + np.where(fin[frm], _BLUR_W[k + 3] * vals[frm], 0.0)
            wsum[dst] = wsum[dst] + np.where(fin[frm], _BLUR_W[k + 3], 0.0)
        with np.errstate(invalid="ignore", divide="ignore"):
            out = np.where(wsum > 0.0, acc / wsum, np.nan).astype(np.float32)
    return out


def restrict(img: np.ndarray):
    """Mean of the finite voxels of each 2x2x2 block (2 or 1 along kept axes), double accumulation in z,y,x order."""
    s, r, c = img.shape
    (cs, cr, cc), _, (hz, hy, hx) = coarsen(img.shape, (1, 1, 1))
    acc = np.zeros((cs, cr, cc), dtype=np.float64)
    cnt = np.zeros((cs, cr, cc), dtype=np.int64)
    for dz in range(2 if hz else 1):
        for dy in range(2 if hy else 1):
            for dx in range(2 if hx else 1):
                sub = img[dz::2 if hz else 1, dy::2 if hy else 1, dx::2 if hx else 1].astype(np.float64)
                fin = np.isfinite(sub)
                a = np.zeros_like(acc)
                k = np.zeros_like(cnt)
                a[:sub.shape[0], :sub.shape[1], :sub.shape[2]] = np.where(fin, sub, 0.0)
                k[:sub.shape[0], :sub.shape[1], :sub.shape[2]] = fin
                acc = acc + a
                cnt = cnt + k
    with np.errstate(invalid="ignore", divide="ignore"):
        out = np.where(cnt > 0, acc / cnt, np.nan)
    return out.astype(np.float32)


def prolong(coarse_field: np.ndarray, fine_shape):
    """Trilinear sampling of the coarse field at u = x/2 - 1/4 (halved axes) clamped to the coarse grid."""
    _, _, (hz, hy, hx) = coarsen(fine_shape, (1, 1, 1))

    def axis(n_f, n_c, halved):
        x = np.arange(n_f, dtype=np.float64)
        u = 0.5 * x - 0.25 if halved else x
        u = np.minimum(np.maximum(u, 0.0), float(n_c - 1))
        f = np.floor(u)
        i0 = f.astype(np.int64)
        return i0, np.minimum(i0 + 1, n_c - 1), u - f

    z0, z1, tz = axis(fine_shape[0], coarse_field.shape[0], hz)
    y0, y1, ty = axis(fine_shape[1], coarse_field.shape[1], hy)
    x0, x1, tx = axis(fine_shape[2], coarse_field.shape[2], hx)
    tz, ty, tx = tz[:, None, None, None], ty[None, :, None, None], tx[None, None, :, None]
    g = lambda zz, yy, xx: coarse_field[zz][:, yy][:, :, xx]
    c00 = g(z0, y0, x0) * (1.0 - tx) + g(z0, y0, x1) * tx
    c10 = g(z0, y1, x0) * (1.0 - tx) + g(z0, y1, x1) * tx
    c01 = g(z1, y0, x0) * (1.0 - tx) + g(z1, y0, x1) * tx
    c11 = g(z1, y1, x0) * (1.0 - tx) + g(z1, y1, x1) * tx
    c0 = c00 * (1.0 - ty) + c10 * ty
    c1 = c01 * (1.0 - ty) + c11 * ty
    return c0 * (1.0 - tz) + c1 * tz


def pyramid_run(fixed, moving, spacing, levels, iters_per_level, sigma_d, sigma_u, diffeo, symmetric):
    """iters_per_level[l] for level l (0 = finest).  Returns (D at the finest level, concatenated MSE history)."""
    F, M, SP = [fixed], [moving], [tuple(spacing)]
    for _ in range(1, levels):
        F.append(restrict(prefilter(F[-1])))
        M.append(restrict(prefilter(M[-1])))
        SP.append(coarsen(F[-2].shape, SP[-1])[1])
    D, hist = None, []
    for l in range(levels - 1, -1, -1):
        D0 = prolong(D, F[l].shape) if D is not None else None
        D, h = run_level(F[l], M[l], SP[l], iters_per_level[l], sigma_d, sigma_u, diffeo, symmetric, D0)
        hist.extend(h.tolist())
    return D, np.array(hist)


# ---- row f1: WarpImages' voxel body for a field and an image on different regular grids ------------------------
# (reference src/Operations/WarpImages.cc:418-456 + src/Alignment_Field.cc:138-153; the trilinear lookups there are
#  Ygor's planar_image_adjacency, which is not vendored: the definition below is ours -- parity unpinned -- and is the
#  one documented in dicomautomaton_b200/csrc/warp_grid.cu)
def _axis(u, n):
    if n == 1:
        ok = np.abs(u) <= 0.5
        z = np.zeros(u.shape, dtype=np.int64)
        return ok, z, z, np.zeros(u.shape)
    ok = (u >= 0.0) & (u <= float(n - 1))
    us = np.where(ok, u, 0.0)
    i0 = np.minimum(np.floor(us).astype(np.int64), n - 2)
    return ok, i0, i0 + 1, us - i0


def _trilinear(data, u):
    """data (S,R,C); u = (ux, uy, uz) index coordinates, arrays of one shape.  Returns (value, ok)."""
    S, R, C = data.shape
    okx, x0, x1, tx = _axis(u[0], C)
    oky, y0, y1, ty = _axis(u[1], R)
    okz, z0, z1, tz = _axis(u[2], S)
    d = data.astype(np.float64)
    with np.errstate(invalid="ignore"):
        c00 = d[z0, y0, x0] * (1.0 - tx) + d[z0, y0, x1] * tx
        c10 = d[z0, y1, x0] * (1.0 - tx) + d[z0, y1, x1] * tx
        c01 = d[z1, y0, x0] * (1.0 - tx) + d[z1, y0, x1] * tx
        c11 = d[z1, y1, x0] * (1.0 - tx) + d[z1, y1, x1] * tx
        c0 = c00 * (1.0 - ty) + c10 * ty
        c1 = c01 * (1.0 - ty) + c11 * ty
        v = c0 * (1.0 - tz) + c1 * tz
    ok = okx & oky & okz & np.isfinite(v)
    return v, ok


def warp_image_on_grid(field, f_spacing, f_origin, image, i_spacing, i_origin, oob_value=0.0, mask=None):
    """Axis-aligned grids (ex, ey, ez = unit axes).  field (S,R,C,3) mm; image (S,R,C) float32."""
    S, R, C = image.shape
    z, y, x = np.meshgrid(np.arange(S, dtype=np.float64), np.arange(R, dtype=np.float64), np.arange(C, dtype=np.float64),
                          indexing="ij")
    p = [i_origin[0] + x * i_spacing[0], i_origin[1] + y * i_spacing[1], i_origin[2] + z * i_spacing[2]]
    uf = [(p[k] - f_origin[k]) / f_spacing[k] for k in range(3)]
    d, okd = [], np.ones(image.shape, dtype=bool)
    for c in range(3):
        v, ok = _trilinear(field[..., c], uf)
        d.append(v)
        okd &= ok
    ui = [(p[k] + np.where(okd, d[k], 0.0) - i_origin[k]) / i_spacing[k] for k in range(3)]
    val, ok = _trilinear(image, ui)
    out = np.where(okd & ok, val, oob_value).astype(np.float32)
    if mask is not None:
        out = np.where(mask.astype(bool), out, image)
    return out


def warp_to_grid(field, f_spacing, f_origin, source, s_spacing, s_origin, placeholder, o_spacing, o_origin, oob_value=0.0,
                 mask=None):
    """The same with the operation's three stacks (src/Operations/WarpImages.cc:385-456), axis-aligned grids, one channel:
    voxels of `placeholder` (on the output grid) selected by `mask` take source(p + D(p)); the others are kept."""
    S, R, C = placeholder.shape
    z, y, x = np.meshgrid(np.arange(S, dtype=np.float64), np.arange(R, dtype=np.float64), np.arange(C, dtype=np.float64),
                          indexing="ij")
    p = [o_origin[0] + x * o_spacing[0], o_origin[1] + y * o_spacing[1], o_origin[2] + z * o_spacing[2]]
    uf = [(p[k] - f_origin[k]) / f_spacing[k] for k in range(3)]
    d, okd = [], np.ones(placeholder.shape, dtype=bool)
    for c in range(3):
        v, ok = _trilinear(field[..., c], uf)
        d.append(v)
        okd &= ok
    us = [(p[k] + np.where(okd, d[k], 0.0) - s_origin[k]) / s_spacing[k] for k in range(3)]
    val, ok = _trilinear(source, us)
    out = np.where(okd & ok, val, oob_value).astype(np.float32)
    if mask is not None:
        out = np.where(mask.astype(bool), out, placeholder.astype(np.float32))
    return out


def virtual_field_v1(kind, shape=(100, 512, 512), bounds=(0.0, 512.0, 0.0, 512.0, 0.0, 100.0)):
    """Row f4: the host loops of GenerateVirtualDataSinusoidalFieldV1.cc:77-193 / GenerateVirtualDataVortexFieldV1.cc:75-178
    restated with numpy (libm sin/cos/exp, as the reference uses).  Returns (field (S,R,C,3), (min, max) magnitude)."""
    S, R, C = shape
    x_min, x_max, y_min, y_max, z_min, z_max = bounds
    px, py, pz = (x_max - x_min) / C, (y_max - y_min) / R, (z_max - z_min) / S
    z, y, x = np.meshgrid(z_min + pz * (np.arange(S) + 0.5), py * np.arange(R), px * np.arange(C), indexing="ij")
    if kind == "sinusoidal_v1":
        wlx, wly, wlz = (x_max - x_min) / 2.0, (y_max - y_min) / 3.0, (z_max - z_min) / 1.5
        pi = np.arccos(-1.0)
        sx, sy, cz = np.sin(2.0 * pi * (x - x_min) / wlx), np.sin(2.0 * pi * (y - y_min) / wly), np.cos(2.0 * pi * (z - z_min) / wlz)
        d = np.stack([sy * cz, sx * cz, sx * sy], axis=-1)
        m = np.sqrt((d * d).sum(-1))
        mn, mx = float(m.min()), float(m.max())
        if mx > mn and mx > 1e-10:
            s = np.where(m > 1e-10, ((m - mn) / (mx - mn)) / np.where(m > 1e-10, m, 1.0), 0.0)
            d = d * s[..., None]
        return d, (mn, mx)
    if kind != "vortex_v1":
        raise ValueError(kind)
    cx, cy, cz0 = 0.5 * (x_min + x_max), 0.5 * (y_min + y_max), 0.5 * (z_min + z_max)
    fx, fy, fz = x - cx, y - cy, z - cz0
    r = np.sqrt(fx * fx + fy * fy)
    ok = r > 1e-6
    rs = np.where(ok, r, 1.0)
    rn = r / min(cx, cy)
    prof = rn * np.exp(-2.0 * (rn - 0.5) * (rn - 0.5))
    zn = np.abs(fz) / ((z_max - z_min) / 2.0)
    zs = 0.3 * np.exp(-zn * zn)
    d = np.stack([np.where(ok, -fy / rs * prof, 0.0), np.where(ok, fx / rs * prof, 0.0),
                  np.where(ok, zs * prof * np.where(fz > 0, 1.0, -1.0), 0.0)], axis=-1)
    m = np.sqrt((d * d).sum(-1))
    mx = float(m.max())
    if mx > 1e-10:
        d = d / mx
    return d, (float(m.min()), mx)


# ---- row a14 / f4: histogram_match, reference src/Alignment_Demons.cc:760-891 (pure std:: code, no Ygor arithmetic) ----
# PINNED: tests/test_oracle_extensions.py checks the reference's own known answers (10, 17.5, 25, 32.5 for a 2x2 ramp
# with 4 bins, constants returned unchanged: src/Alignment_Demons_Tests.cc:204-236).
def histogram_match(source: np.ndarray, reference: np.ndarray, num_bins: int, outlier_fraction: float):
    """Returns (matched float32 array shaped like source, mapped flag)."""
    if source.size == 0 or reference.size == 0:
        raise ValueError("Cannot perform histogram matching: image collection is empty")
    src = source.astype(np.float32)
    sv = np.sort(src[np.isfinite(src)].astype(np.float64))                      # cc:772-798
    rv = np.sort(reference[np.isfinite(reference)].astype(np.float64))
    if sv.size == 0 or rv.size == 0:
        return src.copy(), False                                                  # cc:791-794
    pct = lambda v, p: v[int(p * (v.size - 1))]                                   # cc:801-804
    smin, smax = pct(sv, outlier_fraction), pct(sv, 1.0 - outlier_fraction)
    rmin, rmax = pct(rv, outlier_fraction), pct(rv, 1.0 - outlier_fraction)
    if smax <= smin or rmax <= rmin:
        return src.copy(), False                                                  # cc:812-815

    def cdf(v, lo, hi):                                                           # cc:818-851
        inside = v[(v >= lo) & (v <= hi)]
        b = np.minimum(num_bins - 1, ((inside - lo) / (hi - lo) * num_bins).astype(np.int64))
        c = np.bincount(b, minlength=num_bins).astype(np.float64)
        c = np.cumsum(c)
        return c / c[-1] if c[-1] > 0 else c

    sc, rc = cdf(sv, smin, smax), cdf(rv, rmin, rmax)
    lut = np.empty(num_bins)
    for sb in range(num_bins):                                                    # cc:854-869
        hit = np.nonzero(rc >= sc[sb])[0]
        rb = int(hit[0]) if hit.size else 0
        lut[sb] = rmin + (rmax - rmin) * rb / num_bins
    out = src.copy()
    fin = np.isfinite(src)
    v = src.astype(np.float64)
    with np.errstate(invalid="ignore"):
        b = np.minimum(num_bins - 1, np.where(fin, (v - smin) / (smax - smin) * num_bins, 0.0).astype(np.int64))
    b = np.clip(b, 0, num_bins - 1)
    mapped = np.where(v < smin, rmin, np.where(v > smax, rmax, lut[b]))          # cc:875-886
    out[fin] = mapped[fin].astype(np.float32)
    return out, True


def resample_to_grid(image, s_spacing, s_origin, d_shape, d_spacing, d_origin, oob_value=np.nan):
    """Row f2 (axis-aligned grids): trilinear sample of `image` at the destination voxels' positions (parity unpinned)."""
    S, R, C = d_shape
    z, y, x = np.meshgrid(np.arange(S, dtype=np.float64), np.arange(R, dtype=np.float64), np.arange(C, dtype=np.float64),
                          indexing="ij")
    p = [d_origin[0] + x * d_spacing[0], d_origin[1] + y * d_spacing[1], d_origin[2] + z * d_spacing[2]]
    u = [(p[k] - s_origin[k]) / s_spacing[k] for k in range(3)]
    val, ok = _trilinear(image, u)
    return np.where(ok, val, oob_value).astype(np.float32)
