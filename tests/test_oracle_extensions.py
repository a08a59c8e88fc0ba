"""CPU checks of oracle/extensions.py (the restatement of OUR extension definitions; parity unpinned): with the
extensions switched off it must reduce to the pinned oracle loop bit for bit, and its helpers obey their definitions."""
import numpy as np

from _helpers import phantom_pair, random_smooth_field


def test_level_loop_without_extensions_equals_the_pinned_oracle(oracle_mod):
    from oracle import extensions as ext
    shape, sp = (10, 20, 24), (1.0, 1.2, 2.0)
    F, M = phantom_pair(shape, sp)
    for diffeo in (False, True):
        op = oracle_mod.make_params(max_iterations=6, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                    update_field_smoothing_sigma=1.5, use_diffeomorphic=diffeo)
        Do, ho, _ = oracle_mod.demons_run(F, M, op, sp, kind="port")
        D, h = ext.run_level(F, M, sp, 6, 1.5, 1.5, diffeo, symmetric=False)
        assert np.array_equal(D, Do, equal_nan=True)
        assert np.array_equal(h, ho[:6])


def test_gradient_rule_equals_the_oracle_gradient(oracle_mod):
    from oracle import extensions as ext
    shape, sp = (6, 9, 11), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp)
    F = F.copy()
    F[2, 3, 4] = np.nan
    E = oracle_mod.OracleEngine(F, M, oracle_mod.make_params(max_iterations=1), sp, kind="port")
    try:
        G = E.get("gradient")
    finally:
        E.close()
    assert np.array_equal(ext.gradient_rule(F), G)


def test_symmetric_force_with_equal_gradients_is_the_reference_force(oracle_mod):
    """J = (g + g)/2 = g exactly: the symmetric force of (F, W = F shifted copy with the same gradient) cannot be
    built in general, so check the algebra instead: symmetric_force(F, W, 2*gradF - gradW) equals the oracle's force."""
    from oracle import extensions as ext
    shape, sp = (5, 12, 14), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp)
    E = oracle_mod.OracleEngine(F, M, oracle_mod.make_params(max_iterations=1), sp, kind="port")
    try:
        mse, n = E.compute_update_and_mse()
        U_ref, G = E.get("update"), E.get("gradient")
    finally:
        E.close()
    U, mse2, n2 = ext.symmetric_force(F, M, 2.0 * G - ext.gradient_rule(M))
    assert n2 == n and abs(mse2 - mse) <= 1e-12 * mse
    assert np.abs(U - U_ref).max() <= 1e-12


def test_restrict_prolong_definitions():
    from oracle import extensions as ext
    img = np.arange(4 * 10 * 12, dtype=np.float32).reshape(4, 10, 12)   # 4 slices: z is kept, y and x are halved
    c = ext.restrict(img)
    assert c.shape == (4, 5, 6)
    assert c[1, 2, 3] == np.float32(img[1, 4:6, 6:8].astype(np.float64).mean())
    img[0, 0, 0] = np.nan
    assert ext.restrict(img)[0, 0, 0] == np.float32(np.mean([img[0, 0, 1], img[0, 1, 0], img[0, 1, 1]], dtype=np.float64))
    const = np.full((4, 5, 6, 3), 1.25)
    assert np.array_equal(ext.prolong(const, (4, 10, 12)), np.full((4, 10, 12, 3), 1.25))
    ramp = np.zeros((4, 5, 6, 3))
    ramp[..., 0] = np.arange(6)[None, None, :] * 2.0 + 0.5   # value = coarse-centre position in fine index units
    p = ext.prolong(ramp, (4, 10, 12))
    assert np.allclose(p[2, 3, 1:-1, 0], np.arange(12)[1:-1])  # linear functions are reproduced away from the clamped ends


def test_histogram_match_restatement_reproduces_the_reference_known_answers():
    """src/Alignment_Demons_Tests.cc:204-236: a 2x2 ramp 0..3 matched to 10..40 with 4 bins and no outliers gives exactly
    10, 17.5, 25, 32.5; constant source or constant reference return the source unchanged."""
    from oracle import extensions as ext
    src = np.array([[0, 1], [2, 3]], dtype=np.float32)
    ref = np.array([[10, 20], [30, 40]], dtype=np.float32)
    out, mapped = ext.histogram_match(src, ref, 4, 0.0)
    assert mapped and np.array_equal(out, np.array([[10.0, 17.5], [25.0, 32.5]], dtype=np.float32))
    c5, c10 = np.full((2, 2), 5.0, np.float32), np.full((2, 2), 10.0, np.float32)
    o, m = ext.histogram_match(c5, c10, 8, 0.0)
    assert not m and np.array_equal(o, c5)
    o, m = ext.histogram_match(src, c10, 8, 0.0)
    assert not m and np.array_equal(o, src)


def test_virtual_fields_have_the_properties_the_generator_operations_state():
    """Row f4 restatement (oracle/extensions.py::virtual_field_v1) of GenerateVirtualData{Sinusoidal,Vortex}FieldV1: the
    operations' own descriptions are the known answers -- sinusoidal: 'magnitude ... normalized everywhere to a maximum of
    1.0 and a minimum of 0.0' (SinusoidalFieldV1.cc:33-35); vortex: tangential in the xy plane about the centre, largest
    magnitude 1 (VortexFieldV1.cc:100-178)."""
    from oracle import extensions as ext
    shape, bounds = (20, 96, 96), (0.0, 512.0, 0.0, 512.0, 0.0, 100.0)
    d, (mn, mx) = ext.virtual_field_v1("sinusoidal_v1", shape, bounds)
    m = np.sqrt((d * d).sum(-1))
    assert abs(m.max() - 1.0) < 1e-12 and m.min() < 1e-12 and 0.0 <= mn < mx <= np.sqrt(3.0) + 1e-12
    assert np.all(d[:, 0, 0, :] == 0.0)   # x = y = 0: every component has a sin(0) factor
    v, (vmn, vmx) = ext.virtual_field_v1("vortex_v1", shape, bounds)
    vm = np.sqrt((v * v).sum(-1))
    assert abs(vm.max() - 1.0) < 1e-12
    S, R, C = shape
    y, x = np.meshgrid(512.0 / R * np.arange(R) - 256.0, 512.0 / C * np.arange(C) - 256.0, indexing="ij")
    radial = v[..., 0] * x[None] + v[..., 1] * y[None]
    assert np.abs(radial).max() < 1e-9                      # in-plane displacement is tangential
    assert np.all(v[S // 2:, ..., 2] >= 0.0) and np.all(v[:S // 2, ..., 2] <= 0.0)   # dz points away from the mid-plane
    # the two-pass structure: the second pass is a pure rescale of the first
    full, _ = ext.virtual_field_v1("vortex_v1", (4, 33, 47), bounds)
    assert full.shape == (4, 33, 47, 3) and np.isfinite(full).all()
