"""z-slab sharding (BASELINE config 4) on ONE GPU: several slab engines of one process, halos moved by
device copies, must reproduce the whole-volume engine BIT FOR BIT (same arithmetic per voxel; only the MSE totals are
summed in another order).  The multi-process NCCL transport is exercised by tests/multi_gpu/run_slab_nccl.py under
`gpurun --gpus 2`; its rank logic is the same SlabGroup code."""
import numpy as np
import pytest

from _helpers import phantom_pair

pytestmark = pytest.mark.gpu


def _params(mode, iters, sigma, diffeo, threshold=0.0, **kw):
    from dicomautomaton_b200.demons import AlignViaDemonsParams
    return AlignViaDemonsParams(max_iterations=iters, convergence_threshold=threshold,
                                deformation_field_smoothing_sigma=sigma, update_field_smoothing_sigma=sigma,
                                use_diffeomorphic=diffeo, verbosity=0, field_mode=mode, **kw)


def _whole(F, M, sp, par):
    from dicomautomaton_b200.demons import AlignViaDemons, demons_volume
    out, hist, st = AlignViaDemons(par, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
    assert out is not None
    return out.data, hist, st


def _slabs(F, M, sp, par, n, halo=-1):
    from dicomautomaton_b200.demons import DemonsSlabGroup
    g = DemonsSlabGroup.local(F.shape, sp, par, devices=[0] * n, halo_planes=halo)
    try:
        g.load(F, M)
        hist, st = g.run()
        assert g.owned_range() == (0, F.shape[0])
        D = g.get_deformation()
        hb = g.halo_bytes
    finally:
        g.close()
    return D, hist, st, hb


@pytest.mark.parametrize("mode", ["fp64_strict", "fp64", "mixed", "fp32"])
@pytest.mark.parametrize("diffeo", [False, True])
@pytest.mark.parametrize("n", [2, 3])
def test_slab_group_equals_whole_volume_engine(mode, diffeo, n):
    shape, sp, sigma, iters = (48, 40, 72), (1.0, 1.0, 1.0), 1.5, 10
    F, M = phantom_pair(shape, sp)
    par = _params(mode, iters, sigma, diffeo)
    Dw, hw, stw = _whole(F, M, sp, par)
    Ds, hs, sts, hb = _slabs(F, M, sp, par, n)
    assert sts.iterations_done == stw.iterations_done == iters
    assert np.array_equal(Dw, Ds), f"max diff {np.abs(Dw - Ds).max():.3e}"
    assert np.max(np.abs(hs - hw) / hw) <= (1e-6 if mode in ("fp32", "mixed") else 1e-13)
    assert hb > 0


@pytest.mark.parametrize("mode", ["fp64_strict", "mixed"])
@pytest.mark.parametrize("diffeo", [False, True])
def test_slab_group_with_the_symmetric_force(mode, diffeo):
    """Extension a11 on slabs: the gradient of the warped image needs W one plane beyond the owned range (one plane of the
    final D travels per iteration and W is evaluated on owned +- 1 planes).  Same field as the whole-volume engine, bit for
    bit, and different from the fixed-gradient force."""
    shape, sp, sigma, iters = (45, 40, 72), (1.0, 1.0, 1.0), 1.5, 8
    F, M = phantom_pair(shape, sp)
    par = _params(mode, iters, sigma, diffeo, use_symmetric_force=True)
    Dw, hw, stw = _whole(F, M, sp, par)
    Ds, hs, sts, hb = _slabs(F, M, sp, par, 3)
    assert sts.iterations_done == stw.iterations_done == iters
    assert np.array_equal(Dw, Ds), f"max diff {np.abs(Dw - Ds).max():.3e}"
    assert np.max(np.abs(hs - hw) / hw) <= 1e-6
    Dn, _, _ = _whole(F, M, sp, _params(mode, iters, sigma, diffeo))
    assert np.abs(Dn - Dw).max() > 1e-3


@pytest.mark.parametrize("symmetric", [False, True])
def test_pyramid_with_the_finest_level_on_slabs(monkeypatch, symmetric):
    """Extension a12 on slabs (BASELINE config 3 across GPUs): the coarse levels run whole on one device, the finest level --
    warm-started from the prolonged coarse field, halo planes included -- on a slab group.  Same field as the one-device
    pyramid, bit for bit.  DCMA_B200_SLAB_DEVICES maps both slabs onto device 0 so that a one-GPU box exercises the path;
    the one-shot call with n_devices = 2 (no pyramid) takes the same device list."""
    monkeypatch.setenv("DCMA_B200_SLAB_DEVICES", "0,0")
    shape, sp = (48, 64, 64), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp, amplitude_mm=5.0)
    kw = dict(pyramid_levels=3, pyramid_iterations=(12, 12, 0), use_symmetric_force=symmetric)
    Dw, hw, stw = _whole(F, M, sp, _params("fp64_strict", 8, 1.5, True, **kw))
    Ds, hs, sts = _whole(F, M, sp, _params("fp64_strict", 8, 1.5, True, n_devices=2, **kw))
    assert sts.iterations_done == stw.iterations_done == 32
    assert np.array_equal(Dw, Ds), f"max diff {np.abs(Dw - Ds).max():.3e}"
    assert np.max(np.abs(hs - hw) / hw) <= 1e-12
    assert np.abs(Dw).max() > 3.0   # the warm start is far larger than the default 4-plane halo rule assumes per iteration
    # the plain sharded one-shot call honours the same device list
    D1, _, _ = _whole(F, M, sp, _params("fp64", 6, 1.5, True))
    D2, _, _ = _whole(F, M, sp, _params("fp64", 6, 1.5, True, n_devices=2))
    assert np.array_equal(D1, D2)


def test_slab_group_anisotropic_literal_smoothing_and_ragged_split():
    """per-axis radii (k_smooth_axis path), slices not divisible by the number of slabs, 2 mm slices (r_z = 2)."""
    shape, sp = (37, 21, 19), (0.9, 1.1, 2.0)
    F, M = phantom_pair(shape, sp)
    par = _params("fp64_strict", 8, 1.5, True)
    Dw, hw, _ = _whole(F, M, sp, par)
    Ds, hs, _, _ = _slabs(F, M, sp, par, 3)
    assert np.array_equal(Dw, Ds)


def test_slab_group_stop_rule_is_global():
    """The stop rule sees the MSE of the WHOLE volume: same iteration count and field as the single engine."""
    shape, sp = (40, 32, 36), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp, amplitude_mm=1.0)
    base = _params("fp64", 60, 1.0, False)
    _, h, _ = _whole(F, M, sp, base)
    ch = np.abs(np.diff(h))
    thr = float(0.5 * (ch[19] + ch[20]))
    par = _params("fp64", 60, 1.0, False, threshold=thr, check_every=3)
    Dw, hw, stw = _whole(F, M, sp, par)
    Ds, hs, sts, _ = _slabs(F, M, sp, par, 2)
    assert 1 < stw.iterations_done < 60
    assert sts.iterations_done == stw.iterations_done
    assert np.array_equal(Dw, Ds)


def test_slab_group_reports_a_halo_that_is_too_thin():
    """A composition gather that needs a plane the slab does not hold must fail the run, not return wrong data."""
    from dicomautomaton_b200 import capi
    shape, sp = (48, 24, 32), (1.0, 1.0, 0.25)   # 0.25 mm slices: a 2-voxel update is 8 planes of z displacement
    F, M = phantom_pair(shape, sp, amplitude_mm=3.0)
    par = _params("fp64", 30, 0.3, True, max_update_magnitude=6.0)
    with pytest.raises(capi.DcmaError):
        _slabs(F, M, sp, par, 2, halo=2)


def test_one_shot_call_with_n_devices_uses_slabs():
    """dcma_b200_demons_register with n_devices = 1 device repeated is not possible (devices 0..n-1), so this only
    checks the single-device path is untouched by the new fields and that n_devices = 1 equals n_devices = 0."""
    shape, sp = (16, 24, 28), (1.0, 1.0, 1.0)
    F, M = phantom_pair(shape, sp)
    a, _, _ = _whole(F, M, sp, _params("fp64", 5, 1.5, True))
    b, _, _ = _whole(F, M, sp, _params("fp64", 5, 1.5, True, n_devices=1))
    assert np.array_equal(a, b)
