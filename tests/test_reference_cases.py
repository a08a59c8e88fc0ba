"""The reference's own Demons test cases (src/Alignment_Demons_Tests.cc:301-656), restated at the `demons_volume`
level -- same images, same parameters, same assertions -- and run against (a) the CPU oracle, which shows the oracle
satisfies the reference's tests, and (b) the CUDA path in every field mode (`-m gpu`).

The reference's fixtures are single-slice axis-aligned stacks with unit spacing (`make_test_image_collection`,
:45-80), i.e. exactly a (1, rows, cols) demons_volume.  `warp_image_with_field` is restated with numpy for these 2-D
cases (bilinear, NaN outside: :896-941 with an identity-geometry field lookup).
"""
import numpy as np
import pytest

RUNNERS = ["oracle", pytest.param("fp64_strict", marks=pytest.mark.gpu), pytest.param("fp64", marks=pytest.mark.gpu),
           pytest.param("fp32", marks=pytest.mark.gpu)]


def make_image(rows, cols, fn):
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    return np.asarray(fn(r.astype(np.float64), c.astype(np.float64)), dtype=np.float32).reshape(1, rows, cols)


def align(runner, oracle, moving, fixed, **kw):
    """AlignViaDemons(params, moving, stationary) -> deformation (1, rows, cols, 3) or None."""
    defaults = dict(max_iterations=100, convergence_threshold=0.001, deformation_field_smoothing_sigma=1.0,
                    update_field_smoothing_sigma=0.5, use_diffeomorphic=False, normalization_factor=1.0,
                    max_update_magnitude=2.0)
    defaults.update(kw)
    if runner == "oracle":
        D, hist, n = oracle.demons_run(fixed, moving, oracle.make_params(**defaults), kind="port")
        return D, n
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, demons_volume
    out, hist, st = AlignViaDemons(AlignViaDemonsParams(verbosity=0, field_mode=runner, **defaults),
                                   demons_volume(moving), demons_volume(fixed), return_stats=True)
    return (None, 0) if out is None else (out.data, st.iterations_done)


def warp2d(img, D):
    """warped(p) = img(p + D(p)), bilinear (2-D restatement of warp_image_with_field, :896-941).

    The reference samples through Ygor's planar_image_adjacency, whose in-plane rule treats a pixel as an AREA: a
    position up to half a pixel beyond the outermost centres is still inside and is sampled clamped to the edge;
    further out it is NaN.  (Ygor is not vendored; this rule is the one the reference's own count tolerances in
    :329-376 and :408-447 require.  The engine's own warp_moving is stricter and is pinned separately.)"""
    _, R, C = img.shape
    out = np.full((1, R, C), np.nan, dtype=np.float32)
    for r in range(R):
        for c in range(C):
            sx, sy = c + D[0, r, c, 0], r + D[0, r, c, 1]
            if not (-0.5 <= sx <= C - 0.5 and -0.5 <= sy <= R - 0.5):
                continue
            sx, sy = min(max(sx, 0.0), C - 1.0), min(max(sy, 0.0), R - 1.0)
            x0, y0 = min(int(np.floor(sx)), C - 2), min(int(np.floor(sy)), R - 2)
            tx, ty = sx - x0, sy - y0
            v = (img[0, y0, x0] * (1 - tx) + img[0, y0, x0 + 1] * tx) * (1 - ty) + \
                (img[0, y0 + 1, x0] * (1 - tx) + img[0, y0 + 1, x0 + 1] * tx) * ty
            out[0, r, c] = v
    return out


def mse_and_count(a, b):
    ok = np.isfinite(a) & np.isfinite(b)
    n = int(ok.sum())
    return (float(((a[ok].astype(np.float64) - b[ok]) ** 2).mean()) if n else 0.0), n


def blob(rows, cols, r0, c0, width):
    return make_image(rows, cols, lambda r, c: 100.0 * np.exp(-((r - r0) ** 2 + (c - c0) ** 2) / width))


@pytest.mark.parametrize("runner", RUNNERS)
def test_identical_images_give_zero_field(runner, oracle_mod):
    """:311-327"""
    img = make_image(3, 3, lambda r, c: r + c)
    D, _ = align(runner, oracle_mod, img, img, max_iterations=3, convergence_threshold=0.0,
                 deformation_field_smoothing_sigma=0.0, update_field_smoothing_sigma=0.0)
    assert np.abs(D).max() < 1e-6


@pytest.mark.parametrize("runner", RUNNERS)
@pytest.mark.parametrize("diffeo", [False, True])
def test_improves_mse_for_shifted_ramp(runner, diffeo, oracle_mod):
    """:329-376"""
    rows = cols = 5
    fixed = make_image(rows, cols, lambda r, c: c)
    moving = make_image(rows, cols, lambda r, c: np.minimum(cols - 1, c + 1))
    mse_before, count_before = mse_and_count(fixed, moving)
    D, _ = align(runner, oracle_mod, moving, fixed, max_iterations=15, convergence_threshold=0.0,
                 deformation_field_smoothing_sigma=0.0, update_field_smoothing_sigma=0.0, max_update_magnitude=1.0,
                 use_diffeomorphic=diffeo)
    mse_after, count_after = mse_and_count(fixed, warp2d(moving, D))
    assert count_after >= count_before - max(rows, cols)
    assert mse_after < mse_before


@pytest.mark.parametrize("runner", RUNNERS)
def test_converges_for_gaussian_blob_shift(runner, oracle_mod):
    """:408-447"""
    N = 20
    fixed, moving = blob(N, N, 10.0, 10.0, 8.0), blob(N, N, 10.0, 12.0, 8.0)
    mse_before, count_before = mse_and_count(fixed, moving)
    D, _ = align(runner, oracle_mod, moving, fixed, max_iterations=200, convergence_threshold=0.0,
                 deformation_field_smoothing_sigma=1.0, update_field_smoothing_sigma=0.0)
    warped = warp2d(moving, D)
    mse_after, count_after = mse_and_count(fixed, warped)
    assert mse_after < 0.1 * mse_before
    assert count_after >= count_before - 3 * N
    # :525-579 no drift along y for a pure x shift
    w = np.nan_to_num(warped[0].astype(np.float64))
    f = fixed[0].astype(np.float64)
    rows = np.arange(N)[:, None]
    assert abs((w * rows).sum() / w.sum() - (f * rows).sum() / f.sum()) < 0.1


@pytest.mark.parametrize("runner", RUNNERS)
def test_no_drift_for_identical_blobs(runner, oracle_mod):
    """:450-482"""
    N = 16
    img = blob(N, N, 8.0, 8.0, 6.0)
    D, _ = align(runner, oracle_mod, img, img, max_iterations=200, convergence_threshold=0.0,
                 deformation_field_smoothing_sigma=1.0, update_field_smoothing_sigma=0.0)
    assert np.abs(D).max() < 1e-6
    mse, count = mse_and_count(img, warp2d(img, D))
    assert mse < 1e-6 and count == N * N


@pytest.mark.parametrize("runner", RUNNERS)
def test_mse_decreases_with_more_iterations(runner, oracle_mod):
    """:485-522"""
    N = 16
    fixed, moving = blob(N, N, 8.0, 8.0, 6.0), blob(N, N, 8.0, 9.0, 6.0)
    prev = np.inf
    for iters in (5, 20, 50, 100):
        D, _ = align(runner, oracle_mod, moving, fixed, max_iterations=iters, convergence_threshold=0.0,
                     deformation_field_smoothing_sigma=0.5, update_field_smoothing_sigma=0.0)
        mse, _ = mse_and_count(fixed, warp2d(moving, D))
        assert mse < prev
        prev = mse


@pytest.mark.parametrize("runner", RUNNERS)
def test_diffeomorphic_converges_for_gaussian_blob_shift(runner, oracle_mod):
    """:582-618"""
    N = 20
    fixed, moving = blob(N, N, 10.0, 10.0, 8.0), blob(N, N, 10.0, 12.0, 8.0)
    mse_before, count_before = mse_and_count(fixed, moving)
    D, _ = align(runner, oracle_mod, moving, fixed, max_iterations=200, convergence_threshold=0.0,
                 deformation_field_smoothing_sigma=1.0, update_field_smoothing_sigma=0.5, use_diffeomorphic=True)
    mse_after, count_after = mse_and_count(fixed, warp2d(moving, D))
    assert mse_after < 0.1 * mse_before
    assert count_after >= count_before - 3 * N


@pytest.mark.parametrize("runner", RUNNERS)
def test_convergence_threshold_stops_iteration(runner, oracle_mod):
    """:621-656"""
    N = 16
    fixed, moving = blob(N, N, 8.0, 8.0, 6.0), blob(N, N, 8.0, 9.0, 6.0)
    mse_before, _ = mse_and_count(fixed, moving)
    D, n = align(runner, oracle_mod, moving, fixed, max_iterations=10000, convergence_threshold=1e-3,
                 deformation_field_smoothing_sigma=1.0, update_field_smoothing_sigma=0.0)
    assert n < 10000
    mse_after, _ = mse_and_count(fixed, warp2d(moving, D))
    assert mse_after < 0.5 * mse_before


@pytest.mark.gpu
def test_half_pixel_warp_is_bilinear_not_nearest():
    """:379-405 -- through the CUDA warp entry point."""
    from dicomautomaton_b200.demons import demons_volume, warp_image_with_field
    N = 10
    img = make_image(N, N, lambda r, c: c)
    D = np.zeros((1, N, N, 3))
    D[..., 0] = 0.5
    w = warp_image_with_field(demons_volume(img), D)
    for col in range(1, N - 2):
        assert w[0, 5, col, 0] == pytest.approx(col + 0.5, rel=0.01)


@pytest.mark.gpu
def test_zero_field_warp_semantics():
    """:276-299 restated for the engine's warp_moving: interior voxels are reproduced exactly; the last row/column
    become NaN because corner x0+1 is fetched even when the weight is 0 (cc:508-538)."""
    from dicomautomaton_b200.demons import demons_volume, warp_image_with_field
    img = make_image(6, 7, lambda r, c: 3 * r + c)
    w = warp_image_with_field(demons_volume(img), np.zeros((1, 6, 7, 3)))[..., 0]
    assert np.array_equal(w[0, :-1, :-1], img[0, :-1, :-1])
    assert np.isnan(w[0, -1, :]).all() and np.isnan(w[0, :, -1]).all()
