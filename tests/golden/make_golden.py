#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference engine (oracle/_ref/libdemons_ref.so, built from
/root/reference by oracle/build_ref.py).  Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Each file holds float32 inputs (fixed, moving), the parameters, and the reference's outputs: the fp64 deformation field
after the loop, the MSE history, and -- for stage-level pins -- the update / warped buffers after named stages.  The
inputs are tiny (<= 12x14x16) so the fixtures stay a few hundred kB in total.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from _helpers import phantom_pair, random_smooth_field  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (shape, spacing, params)
    "std_3d": ((10, 12, 14), (1.0, 1.0, 1.0), dict(max_iterations=8, convergence_threshold=0.0,
                                                   deformation_field_smoothing_sigma=1.0, use_diffeomorphic=False)),
    "diffeo_3d": ((10, 12, 14), (1.0, 1.0, 1.0), dict(max_iterations=8, convergence_threshold=0.0,
                                                      deformation_field_smoothing_sigma=1.5,
                                                      update_field_smoothing_sigma=1.5, use_diffeomorphic=True)),
    "diffeo_aniso": ((8, 11, 13), (0.8, 1.2, 2.5), dict(max_iterations=6, convergence_threshold=0.0,
                                                        deformation_field_smoothing_sigma=1.5,
                                                        update_field_smoothing_sigma=0.7, use_diffeomorphic=True)),
    "std_2d": ((1, 14, 16), (1.0, 1.0, 1.0), dict(max_iterations=10, convergence_threshold=0.0,
                                                  deformation_field_smoothing_sigma=1.0, use_diffeomorphic=False)),
    "diffeo_2d": ((1, 14, 16), (1.0, 1.0, 1.0), dict(max_iterations=10, convergence_threshold=0.0,
                                                     deformation_field_smoothing_sigma=1.0,
                                                     update_field_smoothing_sigma=0.5, use_diffeomorphic=True)),
    "converges": ((6, 12, 12), (1.0, 1.0, 1.0), dict(max_iterations=60, convergence_threshold=150.0,
                                                     deformation_field_smoothing_sigma=1.0, use_diffeomorphic=False)),
}


def main():
    if not oracle.available("reference"):
        raise SystemExit("oracle/_ref/libdemons_ref.so is not available (needs /root/reference)")
    for name, (shape, sp, kw) in CASES.items():
        F, M = phantom_pair(shape, sp)
        if name == "std_3d":  # non-finite voxels: the reference's NaN rules are part of the pin
            M = M.copy()
            M[:, :2, :] = np.nan
            F = F.copy()
            F[3, 5, 6] = np.inf
        par = oracle.make_params(**kw)
        D, hist, n = oracle.demons_run(F, M, par, sp, kind="reference")
        out = dict(fixed=F, moving=M, spacing=np.array(sp), deformation=D, mse_history=hist,
                   iterations_done=np.array(n), params=np.array(sorted(kw.items()), dtype=object).astype(str))
        # stage-level pins on the same inputs
        E = oracle.OracleEngine(F, M, par, sp, kind="reference")
        out["gradient"] = E.get("gradient")
        mse, nv = E.compute_update_and_mse()
        out["stage_mse"], out["stage_valid"] = np.array(mse), np.array(nv)
        out["stage_update"] = E.get("update")
        E.smooth_update(kw.get("update_field_smoothing_sigma", 0.5))
        out["stage_update_smoothed"] = E.get("update")
        D0 = random_smooth_field(shape, 2.5, seed=11)
        out["stage_D0"] = D0
        E.set("deformation", D0)
        E.add_update_diffeomorphic()
        out["stage_composed"] = E.get("deformation")
        E.smooth_deformation(kw["deformation_field_smoothing_sigma"])
        out["stage_deformation_smoothed"] = E.get("deformation")
        E.warp_moving()
        out["stage_warped"] = E.get("warped")
        E.close()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} kB, iterations={n}, final mse={hist[-1]:.6g}")


if __name__ == "__main__":
    main()
