"""The two EXTENSIONS (no reference counterpart, "parity unpinned"): symmetric force (row a11) and the multi-resolution
pyramid (row a12).  Checked against oracle/extensions.py -- our own CPU restatement assembled from the pinned oracle
stages -- and by properties: the defaults reproduce the reference path exactly, the pyramid recovers a large warp that
the single-resolution loop does not, and the result then warps a co-registered dose grid (BASELINE config 3)."""
import numpy as np
import pytest

from _helpers import phantom_pair, random_smooth_field

pytestmark = pytest.mark.gpu


def _par(mode, iters, sigma, diffeo, **kw):
    from dicomautomaton_b200.demons import AlignViaDemonsParams
    return AlignViaDemonsParams(max_iterations=iters, convergence_threshold=0.0, deformation_field_smoothing_sigma=sigma,
                                update_field_smoothing_sigma=sigma, use_diffeomorphic=diffeo, verbosity=0, field_mode=mode, **kw)


def _run(par, F, M, sp):
    from dicomautomaton_b200.demons import AlignViaDemons, demons_volume
    out, hist, st = AlignViaDemons(par, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
    assert out is not None
    return out.data, hist, st


@pytest.mark.parametrize("diffeo", [False, True])
@pytest.mark.parametrize("shape,sp", [((14, 24, 28), (1.0, 1.0, 1.0)), ((1, 24, 28), (1.0, 1.0, 1.0)), ((9, 21, 19), (0.9, 1.1, 2.5))])
def test_symmetric_force_matches_the_cpu_restatement(oracle_mod, diffeo, shape, sp):
    from oracle import extensions as ext
    F, M = phantom_pair(shape, sp)
    iters, sigma = 8, 1.5
    Do, ho = ext.run_level(F, M, sp, iters, sigma, sigma, diffeo, symmetric=True)
    for mode, tol in (("fp64_strict", 1e-12), ("fp64", 1e-9), ("fp32", 1e-4)):
        D, h, st = _run(_par(mode, iters, sigma, diffeo, use_symmetric_force=True), F, M, sp)
        assert st.iterations_done == iters
        assert np.array_equal(np.isnan(D), np.isnan(Do))
        assert np.nanmax(np.abs(D - Do)) <= tol, (mode, np.nanmax(np.abs(D - Do)))
        assert np.max(np.abs(h - ho) / ho) <= (1e-4 if mode == "fp32" else 1e-10)


def test_symmetric_force_changes_the_result_and_default_is_off():
    shape, sp = (12, 24, 28), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp)
    a, _, _ = _run(_par("fp64", 6, 1.5, True), F, M, sp)
    b, _, _ = _run(_par("fp64", 6, 1.5, True, use_symmetric_force=False, pyramid_levels=1), F, M, sp)
    c, _, _ = _run(_par("fp64", 6, 1.5, True, use_symmetric_force=True), F, M, sp)
    assert np.array_equal(a, b)
    assert np.abs(a - c).max() > 1e-3


@pytest.mark.parametrize("shape", [(16, 24, 40), (7, 33, 18), (1, 20, 26), (9, 9, 9)])
def test_restrict_and_prolong_match_the_cpu_restatement(shape):
    from dicomautomaton_b200.demons import demons_volume, prolong_field, restrict_image
    from oracle import extensions as ext
    sp = (0.8, 1.0, 2.0)
    F, _ = phantom_pair(shape, sp)
    F = F.copy()
    F[0, 0, :3] = np.nan  # non-finite voxels are left out of the block mean
    c = restrict_image(demons_volume(F, *sp))
    cs, csp, _ = ext.coarsen(shape, sp)
    assert c.data.shape[:3] == cs
    assert (c.pxl_dx, c.pxl_dy, c.pxl_dz) == csp
    assert np.array_equal(c.data[..., 0], ext.restrict(F), equal_nan=True)
    Dc = random_smooth_field(cs, 3.0, seed=2)
    got = prolong_field(Dc, demons_volume(F, *sp))
    assert np.abs(got - ext.prolong(Dc, shape)).max() <= 1e-12


@pytest.mark.parametrize("symmetric", [False, True])
def test_pyramid_matches_the_cpu_restatement(oracle_mod, symmetric):
    from oracle import extensions as ext
    shape, sp = (24, 40, 48), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp, amplitude_mm=4.0)
    iters = (5, 7, 9)  # finest, level 1, level 2
    Do, ho = ext.pyramid_run(F, M, sp, 3, iters, 1.5, 1.5, True, symmetric)
    for mode, tol in (("fp64_strict", 1e-11), ("fp64", 1e-8)):
        D, h, st = _run(_par(mode, iters[0], 1.5, True, pyramid_levels=3, pyramid_iterations=(iters[1], iters[2], 0),
                             use_symmetric_force=symmetric), F, M, sp)
        assert st.iterations_done == sum(iters)
        assert np.nanmax(np.abs(D - Do)) <= tol, (mode, np.nanmax(np.abs(D - Do)))
        assert np.max(np.abs(h - ho) / ho) <= 1e-9


def test_pyramid_recovers_a_large_warp_and_the_field_warps_a_dose_grid():
    """BASELINE config 3 in small: 3-level pyramid + symmetric force on a pair whose warp (6 mm) is beyond the capture
    range of the single-resolution loop, then the result warps a co-registered dose grid (out-of-bounds -> 0.0, the
    WarpImages rule, reference src/Operations/WarpImages.cc:197,452)."""
    import torch
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import demons_volume, warp_image_with_field
    shape, sp = (48, 64, 64), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp, amplitude_mm=6.0)
    single, hs, _ = _run(_par("fp32", 60, 1.5, True), F, M, sp)
    pyr, hp, st = _run(_par("fp32", 20, 1.5, True, pyramid_levels=3, pyramid_iterations=(20, 20, 0), use_symmetric_force=True), F, M, sp)
    assert st.iterations_done == 60
    z = torch.arange(shape[0], dtype=torch.float64).view(-1, 1, 1)
    y = torch.arange(shape[1], dtype=torch.float64).view(1, -1, 1)
    x = torch.arange(shape[2], dtype=torch.float64).view(1, 1, -1)
    w = torch.stack(torch.broadcast_tensors(*phantom.known_warp_mm(x, y, z, shape, amplitude_mm=6.0)), -1).numpy()
    core = (slice(8, -8),) * 3
    err_single = np.sqrt(((single + w)[core] ** 2).mean())
    err_pyr = np.sqrt(((pyr + w)[core] ** 2).mean())
    # better match of the images for the same number of iterations, with a field at least as close to the first-order
    # estimate -w of the known warp (for a 6 mm warp -w is only an approximation of the ideal field D(p) = -w(p + D(p)))
    assert hp[-1] < hs[-1]
    assert err_pyr <= 1.02 * err_single
    zz, yy, xx = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), np.arange(shape[2]), indexing="ij")
    dose = (70.0 * np.exp(-(((zz - 24) / 10.0) ** 2 + ((yy - 32) / 12.0) ** 2 + ((xx - 30) / 14.0) ** 2))).astype(np.float32)
    warped = warp_image_with_field(demons_volume(dose, *sp), pyr, oob_value=0.0)[..., 0]
    assert np.isfinite(warped).all() and warped.min() >= 0.0 and warped.max() <= 70.0 + 1e-3
    assert abs(float(warped.sum()) / float(dose.sum()) - 1.0) < 0.1   # a smooth warp roughly preserves the integral
    assert np.abs(warped - dose).max() > 1.0                           # and it did move the dose


def test_warp_on_another_grid_matches_the_cpu_restatement_and_the_reference_properties():
    """Row f1: the dose grid (coarser, shifted, smaller than the CT grid the field was estimated on).  Checked against
    oracle/extensions.py and against what the reference's own tests pin for this transform: a zero field is the
    identity, a uniform field is a shift (src/Alignment_Field_Tests.cc:291-427), out of bounds gives 0.0
    (src/Operations/WarpImages.cc:197)."""
    from dicomautomaton_b200.demons import make_grid, warp_image_on_grid
    from oracle import extensions as ext
    fshape, fsp, forg = (20, 32, 36), (1.0, 1.0, 2.0), (-10.0, 5.0, 0.0)
    ishape, isp, iorg = (12, 14, 13), (2.5, 2.0, 3.0), (-8.0, 7.5, 1.0)   # partly outside the field's grid
    D = random_smooth_field(fshape, 3.0, seed=7)
    rng = np.random.default_rng(3)
    img = rng.uniform(0.0, 70.0, size=ishape).astype(np.float32)
    fg, ig = make_grid(fshape, fsp, forg), make_grid(ishape, isp, iorg)
    got = warp_image_on_grid(fg, D, ig, img, oob_value=0.0)
    ref = ext.warp_image_on_grid(D, fsp, forg, img, isp, iorg, 0.0)
    assert np.abs(got - ref).max() <= 1e-4 * 70.0 and np.array_equal(got == 0.0, ref == 0.0)
    assert (got == 0.0).any() and (got != 0.0).any()
    # region of interest: voxels outside the mask are left untouched
    mask = np.zeros(ishape, dtype=np.uint8)
    mask[3:9, 4:10, 2:11] = 1
    gm = warp_image_on_grid(fg, D, ig, img, mask=mask, oob_value=0.0)
    assert np.array_equal(gm[mask == 0], img[mask == 0]) and np.array_equal(gm[mask == 1], got[mask == 1])
    # zero field on a grid that covers the image: identity
    big = make_grid((40, 60, 60), (1.0, 1.0, 1.0), (-12.0, 0.0, -2.0))
    ident = warp_image_on_grid(big, np.zeros((40, 60, 60, 3)), ig, img)
    assert np.array_equal(ident, img)
    # uniform field of exactly one image voxel along x: a shift by one column, zeros where the source leaves the image
    U = np.zeros((40, 60, 60, 3))
    U[..., 0] = isp[0]
    sh = warp_image_on_grid(big, U, ig, img)
    assert np.allclose(sh[:, :, :-1], img[:, :, 1:], rtol=0, atol=1e-4) and np.all(sh[:, :, -1] == 0.0)
    # the same grid as the field and an in-grid field: equals the loop's own warp kernel wherever that one is defined
    from dicomautomaton_b200.demons import demons_volume, warp_image_with_field
    F, M = phantom_pair(fshape, (1.0, 1.0, 1.0))
    same = make_grid(fshape, (1.0, 1.0, 1.0))
    a = warp_image_on_grid(same, D, same, M, oob_value=0.0)
    b = warp_image_with_field(demons_volume(M), D, oob_value=0.0)[..., 0]
    inside = b != 0.0
    assert np.abs(a - b)[inside].max() <= 2e-3   # float rounding of the two code paths (HU)


def test_warp_with_three_stacks_as_the_operation_has_them():
    """dcma_b200_warp_to_grid: field on the CT grid, SOURCE stack (ImageSelection) on a dose grid, OUTPUT stack (copy of the
    ReferenceImageSelection) on a third grid; ROI mask and channel selection leave everything else untouched
    (src/Operations/WarpImages.cc:385-456)."""
    from dicomautomaton_b200.demons import make_grid, warp_image_on_grid, warp_to_grid
    from oracle import extensions as ext
    fshape, fsp, forg = (20, 32, 36), (1.0, 1.0, 2.0), (-10.0, 5.0, 0.0)
    sshape, ssp, sorg = (12, 14, 13), (2.5, 2.0, 3.0), (-8.0, 7.5, 1.0)
    oshape, osp, oorg = (9, 17, 21), (1.5, 1.5, 4.0), (-9.0, 6.0, 2.0)
    D = random_smooth_field(fshape, 3.0, seed=9)
    rng = np.random.default_rng(5)
    src = rng.uniform(0.0, 70.0, size=sshape).astype(np.float32)
    place = rng.uniform(-5.0, -1.0, size=oshape).astype(np.float32)  # the reference array's own voxels (negative: recognisable)
    fg, sg, og = make_grid(fshape, fsp, forg), make_grid(sshape, ssp, sorg), make_grid(oshape, osp, oorg)
    mask = np.zeros(oshape, dtype=np.uint8)
    mask[1:8, 2:15, 3:19] = 1
    got = warp_to_grid(fg, D, sg, src, og, place, mask=mask, oob_value=0.0)
    ref = ext.warp_to_grid(D, fsp, forg, src, ssp, sorg, place, osp, oorg, 0.0, mask)
    assert np.abs(got - ref).max() <= 1e-4 * 70.0
    assert np.array_equal(got[mask == 0], place[mask == 0]) and (got[mask == 1] >= 0.0).all()
    assert (got[mask == 1] == 0.0).any() and (got[mask == 1] > 0.0).any()   # some lookups leave a grid, some do not
    # output grid == source grid, no mask: the two-grid call
    same = warp_to_grid(fg, D, sg, src, sg, src, oob_value=0.0)
    assert np.array_equal(same, warp_image_on_grid(fg, D, sg, src, oob_value=0.0))
    # two channels, only channel 1 selected: channel 0 keeps the placeholder's values
    src2 = np.stack([src, src * 0.5], axis=-1)
    place2 = np.stack([place, place], axis=-1)
    sg2, og2 = make_grid(sshape, ssp, sorg, channels=2), make_grid(oshape, osp, oorg, channels=2)
    g2 = warp_to_grid(fg, D, sg2, src2, og2, place2, mask=mask, channel=1, oob_value=0.0)
    assert np.array_equal(g2[..., 0], place) and np.allclose(g2[..., 1][mask == 1], 0.5 * got[mask == 1], rtol=0, atol=1e-4)
    from dicomautomaton_b200 import capi
    with pytest.raises(capi.DcmaError):
        warp_to_grid(fg, D, sg, src, og2, place2, channel=1)   # the source stack has no channel 1


def test_register_stacks_equals_resample_then_register_and_streams_the_field_in_order():
    """dcma_b200_demons_register_stacks (what the C++ drop-in calls for regular stacks): device-side resampling of the moving
    stack + the loop + chunked return of the field == dcma_b200_resample_to_grid followed by dcma_b200_demons_register, bit
    for bit; the chunk notifications tile [0, N) in ascending order."""
    from dicomautomaton_b200.demons import (AlignViaDemons, AlignViaDemonsParams, demons_volume, make_grid, register_stacks,
                                            resample_to_grid)
    fshape, fsp = (140, 128, 132), (1.0, 1.0, 1.5)      # 2.4 M voxels: more than one 2 Mi-voxel chunk
    mshape, msp, morg = (100, 90, 96), (1.4, 1.5, 2.2), (8.0, -2.0, -3.0)   # leaves x < 8 mm of the fixed grid uncovered
    F, _ = phantom_pair(fshape, fsp)
    _, Mraw = phantom_pair(mshape, msp)
    fg, mg = make_grid(fshape, fsp), make_grid(mshape, msp, morg)
    par = AlignViaDemonsParams(max_iterations=6, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                               update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode="fp64_strict")
    chunks = []
    D, hist, st = register_stacks(par, fg, F, mg, Mraw, on_chunk=lambda a, b: chunks.append((a, b)))
    Mres = resample_to_grid(mg, Mraw, fg)
    assert np.isnan(Mres).any() and np.isfinite(Mres).any()   # the moving grid does not cover the fixed one
    ref, href, _ = AlignViaDemons(par, demons_volume(Mres, *fsp), demons_volume(F, *fsp), return_stats=True)
    assert st.iterations_done == 6 and np.array_equal(D, ref.data) and np.array_equal(hist, href)
    n = fshape[0] * fshape[1] * fshape[2]
    assert len(chunks) >= 2 and chunks[0][0] == 0 and chunks[-1][1] == n
    assert all(a < b for a, b in chunks) and all(chunks[i][1] == chunks[i + 1][0] for i in range(len(chunks) - 1))
    # pyramids take the two-call route inside the library; one notification at the end
    chunks.clear()
    par2 = AlignViaDemonsParams(max_iterations=4, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, pyramid_levels=2,
                                pyramid_iterations=(5, 0, 0))
    D2, h2, st2 = register_stacks(par2, fg, F, mg, Mraw, on_chunk=lambda a, b: chunks.append((a, b)))
    ref2, _, _ = AlignViaDemons(par2, demons_volume(Mres, *fsp), demons_volume(F, *fsp), return_stats=True)
    assert chunks == [(0, n)] and st2.iterations_done == 9 and np.array_equal(D2, ref2.data)


@pytest.mark.parametrize("kind", ["sinusoidal_v1", "vortex_v1"])
def test_virtual_field_generators_match_the_host_loops(kind):
    """Row f4: dcma_b200_virtual_field vs the numpy restatement of the generator operations' host loops
    (src/Operations/GenerateVirtualData{Sinusoidal,Vortex}FieldV1.cc), on a reduced grid and on the operations' own
    512 x 512 x 100 grid.  CUDA's sin/cos/exp are not libm's: equal to 1e-13, not bit for bit."""
    from dicomautomaton_b200.demons import virtual_field_v1
    from oracle import extensions as ext
    for shape in ((12, 40, 56), (100, 512, 512)):
        got, grid, (gmn, gmx) = virtual_field_v1(kind, shape)
        ref, (rmn, rmx) = ext.virtual_field_v1(kind, shape)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-13
        assert abs(gmx - rmx) <= 1e-13 and abs(gmn - rmn) <= 1e-13
        assert abs(np.sqrt((got * got).sum(-1)).max() - 1.0) <= 1e-12
        assert (grid.geom.slices, grid.geom.rows, grid.geom.cols) == shape and abs(grid.origin[2] - 0.5 * 100.0 / shape[0]) < 1e-12


def test_generated_field_warps_an_image_and_demons_recovers_it():
    """The use the reference makes of the generators: a known field -> warped image -> registration.  Field from
    dcma_b200_virtual_field (scaled to 3 mm), moving = fixed warped through it (WarpImages' voxel body), then the loop."""
    from dicomautomaton_b200.demons import (AlignViaDemons, AlignViaDemonsParams, demons_volume, make_grid, virtual_field_v1,
                                            warp_image_on_grid)
    shape = (24, 96, 96)
    w, grid, _ = virtual_field_v1("sinusoidal_v1", shape, bounds=(0.0, 96.0, 0.0, 96.0, 0.0, 24.0))
    w *= 3.0
    F, _ = phantom_pair(shape)
    ig = make_grid(shape, (1.0, 1.0, 1.0), origin=tuple(grid.origin))
    M = warp_image_on_grid(grid, w, ig, F, oob_value=float("nan"))
    M = np.where(np.isfinite(M), M, F)
    par = AlignViaDemonsParams(max_iterations=80, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                               update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode="fp64")
    out, hist, st = AlignViaDemons(par, demons_volume(M), demons_volume(F), return_stats=True)
    assert out is not None and st.iterations_done == 80
    assert hist[-1] < 0.2 * hist[0]


def test_gpu_histogram_match_is_bit_identical_to_the_reference_algorithm():
    """dcma_b200_histogram_match vs the pinned restatement of src/Alignment_Demons.cc:760-891: the reference's known
    answers, then random CT-like volumes with NaNs, outliers, different sizes and bin counts -- bit for bit."""
    from dicomautomaton_b200.demons import histogram_match
    from oracle import extensions as ext
    src = np.array([[0, 1], [2, 3]], dtype=np.float32)
    ref = np.array([[10, 20], [30, 40]], dtype=np.float32)
    out, mapped = histogram_match(src, ref, 4, 0.0)
    assert mapped and np.array_equal(out, np.array([[10.0, 17.5], [25.0, 32.5]], dtype=np.float32))
    c10 = np.full((2, 2), 10.0, np.float32)
    o, m = histogram_match(src, c10, 8, 0.0)
    assert not m and np.array_equal(o, src)
    o, m = histogram_match(np.full((2, 2), np.nan, np.float32), ref, 8, 0.0)
    assert not m and np.isnan(o).all()
    rng = np.random.default_rng(11)
    for shape_s, shape_r, bins, frac in (((20, 40, 48), (20, 40, 48), 256, 0.01), ((9, 33, 17), (12, 20, 31), 64, 0.05),
                                         ((6, 50, 50), (6, 50, 50), 5000, 0.0), ((3, 7, 5), (2, 4, 9), 1, 0.0)):
        F, _ = phantom_pair(shape_r)
        _, M = phantom_pair(shape_s, case=2)
        M = (M * 0.8 + 35.0 + rng.normal(0, 5, size=shape_s)).astype(np.float32)
        M[0, 0, :3] = np.nan
        M[1, 2, 3] = np.inf
        F = F.copy()
        F[-1, -1, -1] = -np.inf
        want, wm = ext.histogram_match(M, F, bins, frac)
        got, gm = histogram_match(M, F, bins, frac)
        assert gm == wm
        assert np.array_equal(got, want, equal_nan=True), (shape_s, bins, np.nanmax(np.abs(got - want)))


def test_resample_to_grid_matches_the_restatement_and_the_reference_properties():
    """Row f2: identity on the same grid, NaN outside a smaller source (src/Alignment_Demons_Tests.cc:171-202), and the
    CPU restatement on a shifted, anisotropic destination grid."""
    from dicomautomaton_b200.demons import make_grid, resample_to_grid
    from oracle import extensions as ext
    rng = np.random.default_rng(5)
    img = rng.uniform(-1000.0, 1000.0, size=(7, 9, 11)).astype(np.float32)
    g = make_grid(img.shape, (1.0, 2.0, 3.0), (4.0, -2.0, 1.0))
    assert np.array_equal(resample_to_grid(g, img, g), img)
    bigger = make_grid((9, 11, 13), (1.0, 2.0, 3.0), (3.0, -4.0, -2.0))   # one voxel more on every side
    out = resample_to_grid(g, img, bigger)
    assert np.array_equal(out[1:-1, 1:-1, 1:-1], img) and np.isnan(out[0]).all() and np.isnan(out[:, :, -1]).all()
    dshape, dsp, dorg = (10, 12, 9), (0.7, 1.3, 2.2), (5.1, -1.0, 2.5)
    got = resample_to_grid(g, img, make_grid(dshape, dsp, dorg))
    want = ext.resample_to_grid(img, (1.0, 2.0, 3.0), (4.0, -2.0, 1.0), dshape, dsp, dorg)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.nanmax(np.abs(got - want)) <= 1e-3
