"""CPU tests of the oracle itself (no GPU): the C restatement must equal the compiled reference engine bit for bit,
and both must reproduce the committed golden vectors (tests/golden/*.npz, generated from the reference engine by
tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from _helpers import phantom_pair, random_smooth_field

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def _params_from(oracle, d):
    kw = {k: v for k, v in d["params"]}
    conv = dict(max_iterations=int, convergence_threshold=float, deformation_field_smoothing_sigma=float,
                update_field_smoothing_sigma=float, use_diffeomorphic=lambda s: s == "True",
                normalization_factor=float, max_update_magnitude=float)
    return oracle.make_params(**{k: conv[k](v) for k, v in kw.items()})


def _kinds(oracle):
    kinds = ["port"]
    if oracle.available("reference"):
        kinds.append("reference")
    return kinds


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden_loop(oracle_mod, path):
    d = np.load(path, allow_pickle=True)
    par = _params_from(oracle_mod, d)
    for kind in _kinds(oracle_mod):
        D, hist, n = oracle_mod.demons_run(d["fixed"], d["moving"], par, tuple(d["spacing"]), kind=kind)
        assert n == int(d["iterations_done"])
        assert np.array_equal(D, d["deformation"]), kind
        assert np.array_equal(hist, d["mse_history"]), kind


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden_stages(oracle_mod, path):
    d = np.load(path, allow_pickle=True)
    par = _params_from(oracle_mod, d)
    kw = {k: v for k, v in d["params"]}
    for kind in _kinds(oracle_mod):
        E = oracle_mod.OracleEngine(d["fixed"], d["moving"], par, tuple(d["spacing"]), kind=kind)
        try:
            assert np.array_equal(E.get("gradient"), d["gradient"])
            mse, nv = E.compute_update_and_mse()
            assert mse == float(d["stage_mse"]) and nv == int(d["stage_valid"])
            assert np.array_equal(E.get("update"), d["stage_update"])
            E.smooth_update(float(kw.get("update_field_smoothing_sigma", 0.5)))
            assert np.array_equal(E.get("update"), d["stage_update_smoothed"])
            E.set("deformation", d["stage_D0"])
            E.add_update_diffeomorphic()
            assert np.array_equal(E.get("deformation"), d["stage_composed"])
            E.smooth_deformation(float(kw["deformation_field_smoothing_sigma"]))
            assert np.array_equal(E.get("deformation"), d["stage_deformation_smoothed"])
            E.warp_moving()
            assert np.array_equal(E.get("warped"), d["stage_warped"], equal_nan=True)
        finally:
            E.close()


def _need_ref(oracle):
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref/libdemons_ref.so not available (needs /root/reference or a prebuilt copy)")


@pytest.mark.parametrize("shape,spacing", [((12, 18, 20), (1.0, 1.0, 1.0)), ((1, 20, 24), (1.0, 1.0, 1.0)),
                                           ((9, 13, 11), (0.7, 1.3, 3.0)), ((3, 5, 4), (1.0, 1.0, 1.0))])
@pytest.mark.parametrize("diffeo", [False, True])
def test_port_equals_reference_engine_bit_for_bit(oracle_mod, shape, spacing, diffeo):
    """Pins the C restatement to the reference's own engine on fresh inputs, stage by stage and over a loop."""
    _need_ref(oracle_mod)
    F, M = phantom_pair(shape, spacing)
    F = F.copy()
    M = M.copy()
    M[:, :1, :] = np.nan
    par = oracle_mod.make_params(max_iterations=7, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                 update_field_smoothing_sigma=0.8, use_diffeomorphic=diffeo)
    Dr, hr, nr = oracle_mod.demons_run(F, M, par, spacing, kind="reference")
    Dp, hp, np_ = oracle_mod.demons_run(F, M, par, spacing, kind="port")
    assert nr == np_
    assert np.array_equal(Dr, Dp)
    assert np.array_equal(hr, hp)
    R = oracle_mod.OracleEngine(F, M, par, spacing, kind="reference")
    P = oracle_mod.OracleEngine(F, M, par, spacing, kind="port")
    try:
        D0 = random_smooth_field(shape, 2.0, seed=2)
        for e in (R, P):
            e.compute_update_and_mse()
            e.smooth_update(1.5)
            e.set("deformation", D0)
            e.add_update_diffeomorphic()
            e.smooth_deformation(0.8)
            e.warp_moving()
            e.add_update()
        for buf in ("gradient", "update", "deformation", "warped"):
            assert np.array_equal(R.get(buf), P.get(buf), equal_nan=True), buf
    finally:
        R.close()
        P.close()


def test_port_equals_reference_with_convergence_break(oracle_mod):
    _need_ref(oracle_mod)
    F, M = phantom_pair((8, 20, 20), amplitude_mm=1.0)
    par = oracle_mod.make_params(max_iterations=80, convergence_threshold=1000.0, deformation_field_smoothing_sigma=1.0)
    Dr, hr, nr = oracle_mod.demons_run(F, M, par, kind="reference")
    Dp, hp, np_ = oracle_mod.demons_run(F, M, par, kind="port")
    assert nr == np_ and nr < 80
    assert np.array_equal(Dr, Dp) and np.array_equal(hr, hp)


def test_port_reproduces_the_reference_fixture_of_config_1(oracle_mod):
    """BASELINE config 1 at its real size and length (128^3, standard Demons, sigma_d 1.0 mm, 100 iterations): the C
    restatement must hash to the deformation field the UNMODIFIED reference engine produced
    (tests/golden/full/full_config1_128.npz, written by tests/golden/make_full_size_fixtures.py).  Needs no reference
    sources, so it also pins the port on machines where oracle/_ref cannot be rebuilt."""
    import hashlib
    from dicomautomaton_b200 import phantom
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full", "full_config1_128.npz")
    d = np.load(path, allow_pickle=True)
    shape = tuple(int(v) for v in d["shape"])
    f, m = phantom.make_pair_exact(shape)
    F, M = f.numpy(), m.numpy()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).hexdigest()
    assert sha(F) == str(d["sha_fixed"]) and sha(M) == str(d["sha_moving"])
    conv = dict(max_iterations=int, convergence_threshold=float, deformation_field_smoothing_sigma=float,
                update_field_smoothing_sigma=float, use_diffeomorphic=lambda s: s == "True")
    par = oracle_mod.make_params(**{k: conv[k](v) for k, v in d["params"]})
    D, hist, n = oracle_mod.demons_run(F, M, par, kind="port")
    assert n == int(d["iterations_done"])
    assert sha(D) == str(d["sha_deformation"])
    assert np.array_equal(hist[:n], d["mse_history"])
