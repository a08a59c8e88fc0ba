#!/usr/bin/env python3
"""One-off parity run at BASELINE config 2's FULL size (512x512x256, diffeomorphic, sigma 1.5, N iterations): the
reference engine on the host cores vs the three GPU field modes.  ~2 s of CPU per iteration on 16 cores."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    import oracle
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, demons_volume
    shape = (256, 512, 512)
    f, m = phantom.make_pair(shape, device="cuda")
    F, M = f.cpu().numpy(), m.cpu().numpy()
    kw = dict(max_iterations=iters, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
              update_field_smoothing_sigma=1.5, use_diffeomorphic=True)
    kind = "reference" if oracle.available("reference") else "port"
    t = time.time()
    Do, ho, no = oracle.demons_run(F, M, oracle.make_params(**kw), kind=kind)
    cpu_s = time.time() - t
    rep = {"shape": shape, "iterations": iters, "oracle": kind, "oracle_seconds": cpu_s, "modes": {}}
    for mode in ("fp64_strict", "fp64", "fp32"):
        out, hist, st = AlignViaDemons(AlignViaDemonsParams(verbosity=0, field_mode=mode, **kw), demons_volume(M),
                                       demons_volume(F), return_stats=True)
        err = np.abs(out.data - Do).max(-1)
        rep["modes"][mode] = {"max_abs_dD_mm": float(err.max()), "q9999_mm": float(np.quantile(err[::4, ::4, ::4], 0.9999)),
                              "final_mse_rel": float(abs(hist[-1] - ho[-1]) / ho[-1]), "iterations": int(st.iterations_done),
                              "nan_mask_equal": bool(np.array_equal(np.isnan(out.data), np.isnan(Do))),
                              "loop_ms": float(st.loop_ms)}
        print(mode, rep["modes"][mode], flush=True)
        del out
    print(json.dumps(rep))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rep, open("gpurun_out/full_size_parity.json", "w"), indent=1)


if __name__ == "__main__":
    main()
