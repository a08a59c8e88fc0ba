"""World-size-2 (and 3) gloo tests on CPU of the multi-GPU host logic: case sharding, z-slab partition, the
neighbour halo exchange and the reductions the bench uses (max time, summed units, global MSE terms)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _dist_helpers as parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, slices, halo, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, cols = 5, 7
        full = torch.arange(3 * slices * rows * cols, dtype=torch.float64).reshape(3, slices, rows, cols)
        z0, z1 = parallel.slab_partition(slices, world)[rank]
        ext = parallel.exchange_z_halos(full[:, z0:z1].contiguous(), halo, rank, world)
        lo, hi = max(z0 - halo, 0), min(z1 + halo, slices)
        ok_halo = torch.equal(ext, full[:, lo:hi])
        # reductions
        tmax, usum = parallel.reduce_timing(10.0 * (rank + 1), 100.0 + rank)
        ssq, cnt = parallel.allreduce_mse(float(rank + 1), 10 * (rank + 1))
        ret[rank] = (ok_halo, tmax, usum, ssq, cnt, parallel.shard_cases(7, world, rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,slices,halo", [(2, 16, 4), (2, 9, 3), (3, 20, 4)])
def test_halo_exchange_and_reductions(world, slices, halo):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, slices, halo, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        cases = []
        for r in range(world):
            ok_halo, tmax, usum, ssq, cnt, mine = ret[r]
            assert ok_halo, f"rank {r}: extended slab differs from the whole volume"
            assert tmax == 10.0 * world
            assert usum == sum(100.0 + i for i in range(world))
            assert ssq == sum(range(1, world + 1)) and cnt == 10 * sum(range(1, world + 1))
            cases += mine
        assert sorted(cases) == list(range(7))  # every registration runs exactly once


def test_slab_partition_properties():
    for slices in (1, 7, 256, 1024):
        for world in (1, 2, 3, 4, 8):
            parts = parallel.slab_partition(slices, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == slices
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_radius_rule_matches_reference():
    # max(1, int(3 sigma / pxl)) -- src/Alignment_Demons.cc:574-576
    assert parallel.smoothing_radius(1.5, 1.0) == 4
    assert parallel.smoothing_radius(1.0, 1.0) == 3
    assert parallel.smoothing_radius(0.5, 1.0) == 1
    assert parallel.smoothing_radius(1.5, 2.5) == 1
    assert parallel.smoothing_radius(0.1, 1.0) == 1
    assert parallel.smoothing_radius(0.0, 1.0) == 0


def test_library_slab_partition_matches_python_partition():
    """dcma_b200_slab_partition (pure host arithmetic of the C library, no GPU): owned ranges tile the volume, halos are
    clipped at the global faces, and the owned ranges equal parallel.slab_partition."""
    from dicomautomaton_b200 import capi
    from dicomautomaton_b200.demons import slab_partition
    for slices, world, halo in ((256, 8, 8), (37, 3, 4), (10, 1, 4), (64, 4, 0)):
        ref = parallel.slab_partition(slices, world)
        z = 0
        for r in range(world):
            zoff, local, own0, own1 = slab_partition(slices, world, r, halo)
            assert (zoff + own0, zoff + own1) == ref[r]
            assert zoff == max(0, ref[r][0] - halo) and zoff + local == min(slices, ref[r][1] + halo)
            assert zoff + own0 == z
            z = zoff + own1
        assert z == slices
    with pytest.raises(capi.DcmaError):
        slab_partition(16, 4, 0, 8)   # slabs thinner than the halo
    with pytest.raises(capi.DcmaError):
        slab_partition(3, 4, 0, 0)    # more slabs than slices


# ---- the z-slab exchange SCHEDULE of csrc/slab.cu, emulated on the CPU: two gloo ranks, each running the pinned oracle
# engine on its extended slab, exchanging exactly the planes SlabGroup::run exchanges (r_z planes of U before smooth(U),
# `halo` planes of the smoothed U before the composition, r_z planes of D before smooth(D)) and all-reducing the MSE
# terms of their owned slices.  The owned slices must reproduce the whole-volume oracle: this pins the halo widths and
# the order of the exchanges independently of any GPU (a missing or too thin exchange shows up at the 1e-3 mm level).
# The standard variant is bit-identical.  The diffeomorphic one agrees to the last bits only (<= 1e-12 mm): the oracle
# computes sample coordinates as  z + dz/pxl  with z the LOCAL slice index of its slab, which rounds differently from
# the whole volume's global z -- exactly why the CUDA slab engines carry the global slice index (Geo::zoff) and ARE
# bit-identical to the whole-volume engine (tests/test_gpu_slab.py).
def _put_halos(local, own0, own1, ext, rank, world, n):
    """ext = exchange_z_halos(owned, n): [lower halo?][owned][upper halo?] -> written around the owned slices of `local`."""
    lo = n if rank > 0 else 0
    hi = n if rank < world - 1 else 0
    local[:, own0 - lo:own1 + hi] = ext
    return local


def _slab_worker(rank, world, port, shape, halo, sigma, iters, diffeo, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from dicomautomaton_b200 import phantom
        from dicomautomaton_b200.demons import slab_partition
        oracle.build_port()
        f, m = phantom.make_pair(shape)
        F, M = f.numpy(), m.numpy()
        zoff, local, own0, own1 = slab_partition(shape[0], world, rank, halo)
        Fl, Ml = F[zoff:zoff + local].copy(), M[zoff:zoff + local].copy()
        par = oracle.make_params(max_iterations=iters, convergence_threshold=0.0, deformation_field_smoothing_sigma=sigma,
                                 update_field_smoothing_sigma=sigma, use_diffeomorphic=diffeo)
        E = oracle.OracleEngine(Fl, Ml, par, kind="port")
        rz = parallel.smoothing_radius(sigma, 1.0)
        hist = []

        def exchange(name, n):
            A = torch.from_numpy(np.ascontiguousarray(np.moveaxis(E.get(name), -1, 0)))  # (3, S, R, C)
            ext = parallel.exchange_z_halos(A[:, own0:own1].contiguous(), n, rank, world)
            A = _put_halos(A, own0, own1, ext, rank, world, n)
            E.set(name, np.ascontiguousarray(np.moveaxis(A.numpy(), 0, -1)))

        for _ in range(iters):
            E.compute_update_and_mse()
            fx, wp = E.get("fixed")[own0:own1, ..., 0].astype(np.float64), E.get("warped")[own0:own1, ..., 0].astype(np.float64)
            ok = np.isfinite(fx) & np.isfinite(wp)
            ssq, cnt = parallel.allreduce_mse(float(((fx - wp)[ok] ** 2).sum()), int(ok.sum()))
            hist.append(ssq / cnt)
            if diffeo:
                exchange("update", rz)
                E.smooth_update(sigma)
                exchange("update", halo)
                E.add_update_diffeomorphic()
            else:
                E.add_update()
            exchange("deformation", rz)
            E.smooth_deformation(sigma)
            E.warp_moving()
        ret[rank] = (zoff + own0, zoff + own1, E.get("deformation")[own0:own1].copy(), hist)
        E.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("diffeo", [False, True])
def test_slab_exchange_schedule_reproduces_the_whole_volume_oracle(oracle_mod, diffeo):
    shape, halo, sigma, iters, world = (28, 18, 22), 8, 1.5, 4, 2
    from _helpers import phantom_pair
    F, M = phantom_pair(shape)
    par = oracle_mod.make_params(max_iterations=iters, convergence_threshold=0.0, deformation_field_smoothing_sigma=sigma,
                                 update_field_smoothing_sigma=sigma, use_diffeomorphic=diffeo)
    Dw, hw, _ = oracle_mod.demons_run(F, M, par, kind="port")
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_slab_worker, args=(r, world, port, shape, halo, sigma, iters, diffeo, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        covered = 0
        for r in range(world):
            z0, z1, D, hist = ret[r]
            err = np.abs(D - Dw[z0:z1]).max()
            assert err <= (1e-12 if diffeo else 0.0), f"rank {r}: owned slices differ, max {err:.3e}"
            assert np.allclose(hist, hw[:iters], rtol=1e-12, atol=0)
            covered += z1 - z0
        assert covered == shape[0]
