#!/usr/bin/env python3
"""BASELINE config 3 quality report: 3-level pyramid + symmetric force against single-resolution runs of equal cost, on
the phantom pair with a KNOWN warp w.  Pull convention: warped(p) = moving(p + D(p)) and moving(q) = phantom(q + w(q)), so a
perfect field satisfies  D(p) + w(p + D(p)) = 0;  the reported error is the RMS (and 99th percentile) of that residual in mm
over the voxels at least 8 voxels inside the volume.

    python tools/config3_eval.py [--shape 512,512,512] [--amplitude 4.0] [--texture multiscale|periodic] [--out ...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def residual_stats(D, shape, amplitude, dev):
    import torch
    from dicomautomaton_b200 import phantom
    nz, ny, nx = shape
    b = 8
    acc, cnt, samples = 0.0, 0, []
    for z0 in range(b, nz - b, 16):
        z1 = min(z0 + 16, nz - b)
        d = torch.from_numpy(D[z0:z1, b:ny - b, b:nx - b]).to(dev)
        zs = torch.arange(z0, z1, dtype=torch.float64, device=dev).view(-1, 1, 1)
        ys = torch.arange(b, ny - b, dtype=torch.float64, device=dev).view(1, -1, 1)
        xs = torch.arange(b, nx - b, dtype=torch.float64, device=dev).view(1, 1, -1)
        x, y, z = torch.broadcast_tensors(xs, ys, zs)
        wx, wy, wz = phantom.known_warp_mm(x + d[..., 0], y + d[..., 1], z + d[..., 2], shape, amplitude)
        r = torch.sqrt((d[..., 0] + wx) ** 2 + (d[..., 1] + wy) ** 2 + (d[..., 2] + wz) ** 2)
        acc += float((r * r).sum())
        cnt += r.numel()
        samples.append(r.flatten()[::97].cpu())
    s = torch.cat(samples)
    return {"rms_mm": (acc / cnt) ** 0.5, "p99_mm": float(torch.quantile(s, 0.99)), "max_sampled_mm": float(s.max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,512,512")
    ap.add_argument("--amplitude", type=float, default=4.0)
    ap.add_argument("--mode", default="fp64")
    ap.add_argument("--texture", default="multiscale", choices=["periodic", "multiscale"],
                    help="fine structure of the phantom (dicomautomaton_b200/phantom.py): the benchmark phantom's periodic 7-voxel "
                         "texture is ambiguous beyond ~3.5 voxels of displacement at EVERY pyramid level that resolves it; "
                         "'multiscale' has structure at every scale, which is what a coarse-to-fine scheme feeds on")
    ap.add_argument("--out", default="gpurun_out/config3_quality.json")
    a = ap.parse_args()
    import torch
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, demons_volume
    shape = tuple(int(v) for v in a.shape.split(","))
    dev = torch.device("cuda", 0)
    f, m = phantom.make_pair(shape, amplitude_mm=a.amplitude, device=dev, texture=a.texture)
    F, M = f.cpu().numpy(), m.cpu().numpy()
    del f, m
    iters = (25, 50, 100)  # finest, level 1, level 2
    cost = iters[0] + iters[1] / 8 + iters[2] / 64  # in finest-level iterations
    common = dict(convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5, update_field_smoothing_sigma=1.5,
                  use_diffeomorphic=True, verbosity=0, field_mode=a.mode, device=0)
    runs = {
        "pyramid3_symmetric": dict(max_iterations=iters[0], pyramid_levels=3, pyramid_iterations=(iters[1], iters[2], 0), use_symmetric_force=True),
        "pyramid3_fixed_gradient": dict(max_iterations=iters[0], pyramid_levels=3, pyramid_iterations=(iters[1], iters[2], 0), use_symmetric_force=False),
        "single_symmetric_equal_cost": dict(max_iterations=int(round(cost)), use_symmetric_force=True),
        "single_fixed_gradient_equal_cost": dict(max_iterations=int(round(cost)), use_symmetric_force=False),
        "single_fixed_gradient_200": dict(max_iterations=200, use_symmetric_force=False),
    }
    rep = {"shape": list(shape), "amplitude_mm": a.amplitude, "mode": a.mode, "texture": a.texture,
           "zero_field": residual_stats(np.zeros(shape + (3,)), shape, a.amplitude, dev),
           "pyramid_cost_in_finest_iterations": cost, "runs": {}}
    for name, kw in runs.items():
        par = AlignViaDemonsParams(**common, **kw)
        t0 = time.perf_counter()
        out, hist, st = AlignViaDemons(par, demons_volume(M), demons_volume(F), return_stats=True)
        dt = time.perf_counter() - t0
        r = {"iterations": int(st.iterations_done), "loop_ms": st.loop_ms, "call_s": dt, "mse_first": float(hist[0]), "mse_last": float(hist[-1])}
        if kw.get("pyramid_levels", 1) == 3:
            r["mse_per_level_first_last"] = {"level2": [float(hist[0]), float(hist[iters[2] - 1])],
                                             "level1": [float(hist[iters[2]]), float(hist[iters[2] + iters[1] - 1])],
                                             "level0": [float(hist[iters[2] + iters[1]]), float(hist[-1])]}
        r.update(residual_stats(out.data, shape, a.amplitude, dev))
        rep["runs"][name] = r
        print(name, json.dumps(r), flush=True)
        del out
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
