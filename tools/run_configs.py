#!/usr/bin/env python3
"""Run BASELINE.json's five configurations end to end and record what was measured (not a bench line; bench.py owns the
headline metric on config 2).  Single process unless launched under torchrun (configs 4 and 5 then use all ranks).

    python tools/run_configs.py [--configs 1,2,3,4,5] [--mode fp64] [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,5")
    ap.add_argument("--mode", default="fp64")
    ap.add_argument("--quick", action="store_true", help="fewer iterations / smaller batch (smoke)")
    ap.add_argument("--out", default="gpurun_out/configs.json")
    a = ap.parse_args()
    import torch
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import (AlignViaDemons, AlignViaDemonsParams, DemonsEngine, DemonsSlabGroup,
                                            demons_volume, warp_image_with_field)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    want = {int(c) for c in a.configs.split(",")}
    rep = {"mode": a.mode, "world": world, "gpu": torch.cuda.get_device_name(local), "configs": {}}

    def engine_run(shape, par, case=0):
        f, m = phantom.make_pair(shape, device=dev, case=case)
        eng = DemonsEngine.empty(shape, (1.0, 1.0, 1.0), par)
        eng.load_device(f.data_ptr(), m.data_ptr())
        hist, st = eng.run(par.max_iterations)
        eng.close()
        return hist, st

    if 1 in want and rank == 0:  # 128^3, single-resolution standard Demons, 100 iterations
        par = AlignViaDemonsParams(max_iterations=100, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.0,
                                   use_diffeomorphic=False, verbosity=0, field_mode=a.mode, device=local)
        engine_run((128, 128, 128), par)
        hist, st = engine_run((128, 128, 128), par)
        rep["configs"]["1"] = {"shape": [128, 128, 128], "variant": "standard, sigma_d 1.0 mm (r=3)", "iterations": int(st.iterations_done),
                               "loop_ms": st.loop_ms, "voxel_updates_per_s": 128 ** 3 * st.iterations_done / (st.loop_ms * 1e-3),
                               "mse_first": float(hist[0]), "mse_last": float(hist[-1]), "launches": int(st.kernel_launches)}
    if 2 in want and rank == 0:  # 512x512x256 diffeomorphic, sigma 1.5, 200 iterations
        it = 50 if a.quick else 200
        par = AlignViaDemonsParams(max_iterations=it, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                   update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode=a.mode, device=local)
        hist, st = engine_run((256, 512, 512), par)
        rep["configs"]["2"] = {"shape": [256, 512, 512], "variant": "diffeomorphic, sigma 1.5 mm (r=4)", "iterations": int(st.iterations_done),
                               "loop_ms": st.loop_ms, "voxel_updates_per_s": 256 * 512 * 512 * st.iterations_done / (st.loop_ms * 1e-3),
                               "mse_first": float(hist[0]), "mse_last": float(hist[-1])}
    if 3 in want and rank == 0:  # 3-level pyramid, symmetric force, 512^3, then warp of a co-registered dose grid
        shape = (256, 256, 256) if a.quick else (512, 512, 512)
        iters = (25, 50, 100)  # finest, level 1, level 2 (coarsest first in time)
        # structure at every scale (see tools/config3_eval.py and DESIGN.md 5.4 for why the benchmark phantom's periodic
        # texture cannot show what a pyramid does)
        f, m = phantom.make_pair(shape, amplitude_mm=4.0, device=dev, texture="multiscale")
        F, M = f.cpu().numpy(), m.cpu().numpy()
        par = AlignViaDemonsParams(max_iterations=iters[0], convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                   update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode=a.mode,
                                   device=local, pyramid_levels=3, pyramid_iterations=(iters[1], iters[2], 0), use_symmetric_force=True)
        t0 = time.perf_counter()
        out, hist, st = AlignViaDemons(par, demons_volume(M), demons_volume(F), return_stats=True)
        t_reg = time.perf_counter() - t0
        zz, yy, xx = np.meshgrid(*[np.arange(n, dtype=np.float32) for n in shape], indexing="ij")
        dose = (70.0 * np.exp(-(((zz - shape[0] / 2) / (shape[0] / 5)) ** 2 + ((yy - shape[1] / 2) / (shape[1] / 6)) ** 2 +
                                ((xx - shape[2] * 0.45) / (shape[2] / 5)) ** 2))).astype(np.float32)
        del zz, yy, xx
        t0 = time.perf_counter()
        wd = warp_image_with_field(demons_volume(dose), out.data, oob_value=0.0, device=local)
        t_warp = time.perf_counter() - t0
        rep["configs"]["3"] = {"shape": list(shape), "variant": "3-level pyramid + symmetric force, diffeomorphic, sigma 1.5 mm; iterations coarse->fine 100/50/25; multiscale phantom, 4 mm warp",
                               "iterations": int(st.iterations_done), "loop_ms_all_levels": st.loop_ms, "register_call_s": t_reg,
                               "mse_first_coarsest": float(hist[0]), "mse_first_finest": float(hist[iters[1] + iters[2]]), "mse_last": float(hist[-1]),
                               "dose_warp_call_s": t_warp, "dose_integral_ratio": float(wd.sum() / dose.sum()), "dose_max": float(wd.max())}
        del out, wd, dose
    if 4 in want:  # 1024^3 z-slab sharded (all ranks), 20 iterations
        shape = (512, 512, 512) if a.quick else (1024, 1024, 1024)
        par = AlignViaDemonsParams(max_iterations=20, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                   update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode=a.mode, device=local)
        f, m = phantom.make_pair(shape, device=dev)
        ident = [DemonsSlabGroup.unique_id() if rank == 0 else None]
        if dist is not None:
            dist.broadcast_object_list(ident, src=0)
            grp = DemonsSlabGroup.nccl(shape, (1.0, 1.0, 1.0), par, ident[0], rank, world, local)
        else:
            grp = DemonsSlabGroup.local(shape, (1.0, 1.0, 1.0), par, devices=[local])
        grp.load_device(f.data_ptr(), m.data_ptr())
        grp.run(3)
        grp.load_device(f.data_ptr(), m.data_ptr())
        hist, st = grp.run(20)
        ms = torch.tensor([st.loop_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            n = shape[0] * shape[1] * shape[2]
            rep["configs"]["4"] = {"shape": list(shape), "variant": f"z-slabs over {world} GPU(s), NCCL halos, diffeomorphic sigma 1.5 mm", "iterations": 20,
                                   "loop_ms": float(ms), "voxel_updates_per_s": n * 20 / (float(ms) * 1e-3), "halo_bytes_rank0": grp.halo_bytes,
                                   "mse_first": float(hist[0]), "mse_last": float(hist[-1])}
        grp.close()
        del f, m
    if 5 in want:  # 64 independent 256^3 registrations, round-robin over the ranks, 100 iterations each
        n_cases = 8 if a.quick else 64
        it = 100
        par = AlignViaDemonsParams(max_iterations=it, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                   update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode=a.mode, device=local)
        mine = list(range(rank, n_cases, world))  # round-robin (config 5)
        eng = DemonsEngine.empty((256, 256, 256), (1.0, 1.0, 1.0), par)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        dev_ms, last = 0.0, []
        for c in mine:
            f, m = phantom.make_pair((256, 256, 256), device=dev, case=c)
            eng.load_device(f.data_ptr(), m.data_ptr())
            hist, st = eng.run(it)
            dev_ms += st.loop_ms
            last.append(float(hist[-1]))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        eng.close()
        t = torch.tensor([wall, dev_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            rep["configs"]["5"] = {"cases": n_cases, "shape": [256, 256, 256], "iterations_each": it, "ranks": world,
                                   "wall_s_max_over_ranks_incl_phantom_generation": float(t[0]), "loop_ms_max_over_ranks": float(t[1]),
                                   "voxel_updates_per_s_loop_only": n_cases * 256 ** 3 * it / (float(t[1]) * 1e-3),
                                   "final_mse_rank0_cases": last[:4]}
    if rank == 0:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(rep, open(a.out, "w"), indent=1)
        print(json.dumps(rep, indent=1))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
