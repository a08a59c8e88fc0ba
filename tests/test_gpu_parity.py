"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star: max |d displacement| <= 1e-3 mm, final MSE within 1e-4 relative):
  * field_mode fp64_strict : BIT-IDENTICAL deformation/update/warped (NaN masks included) -- stage by stage and loops
  * field_mode fp64        : <= 1e-9 mm   (FMA contraction and reciprocal multiplies only)
  * field_mode fp32        : <= 1e-4 mm on these sizes (fp32 storage/arithmetic of the fields)
  * field_mode mixed       : <= 1e-5 mm (D and all arithmetic fp64; only U is stored and smoothed as float)
MSE values: the GPU reduces in a tree, the reference serially => 1e-12 relative is the bar for fp64 modes.
"""
import numpy as np
import pytest

from _helpers import phantom_pair, random_smooth_field, ref_kind

pytestmark = pytest.mark.gpu

TOL_MM = {"fp64_strict": 0.0, "fp64": 1e-9, "fp32": 1e-4, "mixed": 1e-5}
TOL_MSE = {"fp64_strict": 1e-12, "fp64": 1e-12, "fp32": 1e-4, "mixed": 1e-6}
# warped intensities (HU): fp32 fields move the sample point by ~1e-7 voxel
TOL_HU = {"fp64_strict": 0.0, "fp64": 1e-9, "fp32": 2e-3, "mixed": 1e-6}



CASES = [
    # (shape (slices, rows, cols), spacing (dx, dy, dz), sigma)
    ((20, 24, 28), (1.0, 1.0, 1.0), 1.5),    # vector path (cols % 4 == 0), r = 4
    ((1, 24, 28), (1.0, 1.0, 1.0), 1.5),     # 2-D branch (slices == 1)
    ((12, 21, 19), (0.9, 1.1, 2.5), 1.5),    # odd extents, anisotropic spacing -> per-axis radii (literal passes)
    ((9, 70, 130), (1.0, 1.0, 1.0), 1.0),    # several tiles in x and y, ragged edges, r = 3
    ((40, 33, 66), (2.0, 2.0, 2.0), 1.5),    # r = 2, cols % 4 == 2 (16-byte path for fp64 only)
    ((6, 5, 7), (1.0, 1.0, 1.0), 0.5),       # tiny: extents smaller than the kernel support, r = 1
]


def _check(name, got, ref, mode, tol=None):
    tol = TOL_MM[mode] if tol is None else tol
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{name}: NaN masks differ"
    if tol == 0.0:
        assert np.array_equal(got, ref, equal_nan=True), \
            f"{name}: not bit-identical, max diff {np.nanmax(np.abs(got - ref)):.3e}"
    else:
        err = np.nanmax(np.abs(np.nan_to_num(got) - np.nan_to_num(ref)))
        assert err <= tol, f"{name}: max abs diff {err:.3e} > {tol:.1e}"


def _params(oracle, mode, iters, sigma, diffeo, fused=True, threshold=0.0, **kw):
    from dicomautomaton_b200.demons import AlignViaDemonsParams
    common = dict(max_iterations=iters, convergence_threshold=threshold, deformation_field_smoothing_sigma=sigma,
                  update_field_smoothing_sigma=sigma, use_diffeomorphic=diffeo)
    common.update(kw)
    return oracle.make_params(**common), AlignViaDemonsParams(verbosity=0, field_mode=mode, fused=fused, **common)


@pytest.mark.parametrize("mode", ["fp64_strict", "fp64", "mixed", "fp32"])
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}" for c in CASES])
def test_stage_level_parity(oracle_mod, mode, fused, case):
    """Each reference engine method in isolation, on identical buffers (oracle buffers are uploaded to the GPU)."""
    from dicomautomaton_b200.demons import DemonsEngine, demons_volume
    shape, sp, sigma = case
    F, M = phantom_pair(shape, sp)
    op, gp = _params(oracle_mod, mode, 1, sigma, True, fused)
    O = oracle_mod.OracleEngine(F, M, op, sp, kind=ref_kind(oracle_mod))
    G = DemonsEngine(demons_volume(F, *sp), demons_volume(M, *sp), gp)
    try:
        _check("gradient", G.get("gradient"), O.get("gradient"), mode, tol=0.0)  # fp64 from float inputs: always exact
        # force + MSE on W = M
        mo, no = O.compute_update_and_mse()
        mg, ng = G.compute_update_and_mse()
        assert ng == no
        assert abs(mg - mo) <= TOL_MSE[mode] * abs(mo)
        _check("update", G.get("update"), O.get("update"), mode)
        # smoothing of the update field
        G.set("update", O.get("update"))
        O.smooth_update(sigma)
        G.smooth_update(sigma)
        _check("smooth_update", G.get("update"), O.get("update"), mode)
        # compose and add on a non-trivial deformation
        D0 = random_smooth_field(shape, 3.0, seed=1)
        for stage in ("add_update_diffeomorphic", "add_update"):
            O.set("deformation", D0)
            G.set("deformation", D0)
            G.set("update", O.get("update"))
            getattr(O, stage)()
            getattr(G, stage)()
            _check(stage, G.get("deformation"), O.get("deformation"), mode)
        G.set("deformation", O.get("deformation"))
        O.smooth_deformation(sigma)
        G.smooth_deformation(sigma)
        _check("smooth_deformation", G.get("deformation"), O.get("deformation"), mode)
        # warp: knife-edge validity at the faces must match (NaN mask is part of _check)
        G.set("deformation", O.get("deformation"))
        O.warp_moving()
        G.warp_moving()
        _check("warped", G.get("warped"), O.get("warped"), mode, tol=TOL_HU[mode])
        # force on the warped image (NaN voxels present now)
        G.set("warped", O.get("warped"))
        mo, no = O.compute_update_and_mse()
        mg, ng = G.compute_update_and_mse()
        assert ng == no and abs(mg - mo) <= TOL_MSE[mode] * abs(mo)
        _check("update(after warp)", G.get("update"), O.get("update"), mode)
    finally:
        G.close()
        O.close()


@pytest.mark.parametrize("mode", ["fp64_strict", "fp64", "mixed", "fp32"])
@pytest.mark.parametrize("diffeo", [False, True])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}" for c in CASES])
def test_loop_parity(oracle_mod, mode, diffeo, case):
    """Whole registrations through the one-shot C-ABI call vs the oracle's AlignViaDemons loop."""
    from dicomautomaton_b200.demons import AlignViaDemons, demons_volume
    shape, sp, sigma = case
    iters = 12
    F, M = phantom_pair(shape, sp)
    op, gp = _params(oracle_mod, mode, iters, sigma, diffeo)
    Do, ho, no = oracle_mod.demons_run(F, M, op, sp, kind=ref_kind(oracle_mod))
    out, hg, st = AlignViaDemons(gp, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
    assert out is not None
    assert st.iterations_done == no == iters
    _check("deformation", out.data, Do, mode)
    assert np.max(np.abs(hg - ho) / np.abs(ho)) <= TOL_MSE[mode]


@pytest.mark.parametrize("mode", ["fp64_strict", "fp64"])
def test_unfused_equals_fused(oracle_mod, mode):
    """fused=0 (one kernel per reference stage, literal smoothing passes) and fused=1 give the same field."""
    from dicomautomaton_b200.demons import AlignViaDemons, demons_volume
    shape, sp = (16, 40, 72), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp)
    outs = []
    for fused in (True, False):
        _, gp = _params(oracle_mod, mode, 10, 1.5, True, fused)
        out = AlignViaDemons(gp, demons_volume(M, *sp), demons_volume(F, *sp))
        outs.append(out.data)
    _check("fused vs unfused", outs[0], outs[1], mode, tol=0.0 if mode == "fp64_strict" else 1e-12)


@pytest.mark.parametrize("diffeo", [False, True], ids=["standard", "diffeomorphic"])
@pytest.mark.parametrize("mode", ["fp64_strict", "fp64", "mixed", "fp32"])
def test_convergence_rule_matches_reference(oracle_mod, mode, diffeo):
    """The stop rule (cc:987-993) runs on the device: same number of iterations, same field as the oracle.

    The host enqueues iterations ahead of the device-side stop rule and polls every `check_every` iterations, so
    1..check_every-1 iterations can be enqueued AFTER the converging one and are skipped on the device.  Each skipped
    iteration still rotates the field pointers on the host (one swap per smoothing call: period 2 for standard Demons,
    period 3 for diffeomorphic with both sigmas), so the cadences below are chosen to give skipped counts that are odd
    and not multiples of 3 -- the returned field must be the one of the converging iteration (cc:987-997)."""
    from dicomautomaton_b200.demons import AlignViaDemons, demons_volume
    shape, sp = (10, 32, 36), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp, amplitude_mm=1.0)
    # pick a threshold that stops the oracle somewhere in the middle of the run
    op, _ = _params(oracle_mod, mode, 60, 1.0, diffeo)
    _, h, _ = oracle_mod.demons_run(F, M, op, sp, kind=ref_kind(oracle_mod))
    changes = np.abs(np.diff(h))
    k = 20
    thr = float(0.5 * (changes[k - 1] + changes[k]))  # |mse[k]-mse[k+1]| < thr first happens near iteration k+1
    op, gp = _params(oracle_mod, mode, 60, 1.0, diffeo, threshold=thr)
    Do, ho, no = oracle_mod.demons_run(F, M, op, sp, kind=ref_kind(oracle_mod))
    assert 1 < no < 60
    skipped_seen = set()
    for check_every in (1, 2, 3, 4, 5, 6, 7, 11):
        gp.check_every = check_every
        # iterations enqueued when the first poll at or after the converging one happens (or the loop ends)
        enq = min(60, -(-no // check_every) * check_every)
        skipped_seen.add(enq - no)
        out, hg, st = AlignViaDemons(gp, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
        assert st.iterations_done == no, f"check_every={check_every}"
        _check(f"deformation (check_every={check_every}, {enq - no} skipped)", out.data, Do, mode)
    # the point of this test: an odd skipped count and one that is not a multiple of 3 have both been exercised
    assert any(s % 2 == 1 for s in skipped_seen) and any(s % 3 != 0 for s in skipped_seen), skipped_seen


@pytest.mark.parametrize("mode", ["fp64_strict", "fp64", "mixed", "fp32"])
def test_non_finite_inputs(oracle_mod, mode):
    """NaN/Inf voxels in fixed and moving (e.g. outside a resampled field of view): same NaN handling as the
    reference -- zero force, excluded from the MSE, zero gradient next to them (cc:238, 297-304, 535-539)."""
    from dicomautomaton_b200.demons import AlignViaDemons, demons_volume
    shape, sp = (8, 24, 32), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp)
    F = F.copy()
    M = M.copy()
    M[:, :3, :] = np.nan        # a band outside the moving field of view
    M[2, 10, 11] = np.inf
    F[5, 7:9, 20:23] = np.nan
    F[1, 1, 1] = -np.inf
    for diffeo in (False, True):
        op, gp = _params(oracle_mod, mode, 8, 1.5, diffeo)
        Do, ho, no = oracle_mod.demons_run(F, M, op, sp, kind=ref_kind(oracle_mod))
        out, hg, st = AlignViaDemons(gp, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
        assert st.iterations_done == no
        _check("deformation", out.data, Do, mode)
        assert np.max(np.abs(hg - ho) / np.abs(ho)) <= TOL_MSE[mode]


@pytest.mark.parametrize("mode", ["fp64_strict", "fp64"])
def test_non_finite_field_values_take_the_literal_smoothing_path(oracle_mod, mode):
    """NaN/Inf inside the FIELD being smoothed: taps are skipped and weights renormalised (cc:626, 655, 684)."""
    from dicomautomaton_b200.demons import DemonsEngine, demons_volume
    shape, sp, sigma = (12, 40, 72), (1.0, 1.0, 1.0), 1.5
    F, M = phantom_pair(shape, sp)
    op, gp = _params(oracle_mod, mode, 1, sigma, True)
    U = random_smooth_field(shape, 1.0, seed=3)
    U[3, 5, 6, 0] = np.nan
    U[0, 0, 0, 1] = np.inf
    U[11, 39, 71, 2] = -np.inf
    U[6, 20:31, 30:41, 1] = np.nan  # a block larger than the kernel support in x and y
    O = oracle_mod.OracleEngine(F, M, op, sp, kind=ref_kind(oracle_mod))
    G = DemonsEngine(demons_volume(F, *sp), demons_volume(M, *sp), gp)
    try:
        O.set("update", U)
        G.set("update", U)
        O.smooth_update(sigma)
        G.smooth_update(sigma)
        _check("smooth with non-finite taps", G.get("update"), O.get("update"), mode)
    finally:
        G.close()
        O.close()


@pytest.mark.parametrize("mode", ["fp64_strict", "fp32"])
def test_multichannel_images(oracle_mod, mode):
    """channels > 1: only channel 0 drives the force (cc:292-295) but every channel is warped (cc:515)."""
    from dicomautomaton_b200.demons import DemonsEngine, demons_volume
    shape, sp = (6, 20, 24), (1.0, 1.0, 1.0)
    F1, M1 = phantom_pair(shape, sp)
    F2, M2 = phantom_pair(shape, sp, case=3)
    F = np.stack([F1, F2], axis=-1)
    M = np.stack([M1, M2], axis=-1)
    op, gp = _params(oracle_mod, mode, 1, 1.0, False)
    O = oracle_mod.OracleEngine(F, M, op, sp, kind=ref_kind(oracle_mod))
    G = DemonsEngine(demons_volume(F, *sp), demons_volume(M, *sp), gp)
    try:
        for _ in range(3):
            mo = O.iterate()
        hist, st = G.run(3)
        assert abs(hist[-1] - mo) <= TOL_MSE[mode] * abs(mo)
        _check("deformation", G.get("deformation"), O.get("deformation"), mode)
        _check("warped (2 channels)", G.get("warped"), O.get("warped"), mode, tol=TOL_HU[mode])
    finally:
        G.close()
        O.close()


def test_warp_image_entry_point(oracle_mod):
    """dcma_b200_warp_image == warp_moving semantics; oob_value replaces NaN (WarpImages uses 0.0)."""
    from dicomautomaton_b200.demons import demons_volume, warp_image_with_field
    shape, sp = (7, 18, 22), (1.0, 1.5, 2.0)
    F, M = phantom_pair(shape, sp)
    D = random_smooth_field(shape, 2.5, seed=5)
    O = oracle_mod.OracleEngine(F, M, oracle_mod.make_params(max_iterations=1), sp, kind=ref_kind(oracle_mod))
    try:
        O.set("deformation", D)
        O.warp_moving()
        ref = O.get("warped")
    finally:
        O.close()
    got = warp_image_with_field(demons_volume(M, *sp), D)
    assert np.array_equal(got, ref, equal_nan=True)
    got0 = warp_image_with_field(demons_volume(M, *sp), D, oob_value=0.0)
    assert np.array_equal(got0, np.where(np.isnan(ref), 0.0, ref).astype(np.float32))


def test_engine_blocks_are_reused_through_the_device_pool():
    """Engines give their device blocks back to a process-lifetime pool (csrc/device_pool.h) and the next engine of the same
    shape takes them from there: no cudaMalloc / cudaFree of several GB per registration.  Results do not depend on it
    (a reused block holds the previous run's bytes; every buffer is written before it is read)."""
    from dicomautomaton_b200 import capi
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, demons_volume
    lib = capi.lib()
    shape, sp = (24, 40, 56), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp)
    par = AlignViaDemonsParams(max_iterations=6, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                               update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode="fp64_strict")
    lib.dcma_b200_device_pool_trim()
    assert lib.dcma_b200_device_pool_bytes() == 0
    a, ha, _ = AlignViaDemons(par, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
    cached = lib.dcma_b200_device_pool_bytes()
    assert cached >= 3 * 3 * 8 * F.size     # at least the three fp64 field allocations
    # another registration in between with other inputs dirties the cached blocks
    AlignViaDemons(par, demons_volume(F, *sp), demons_volume(M[::-1].copy(), *sp))
    b, hb, _ = AlignViaDemons(par, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
    assert lib.dcma_b200_device_pool_bytes() == cached      # same blocks went out and came back
    assert np.array_equal(a.data, b.data) and np.array_equal(ha, hb)
    lib.dcma_b200_device_pool_trim()
    assert lib.dcma_b200_device_pool_bytes() == 0
