#!/usr/bin/env python3
"""ONE process driving several GPUs (the layout of the single-process reference host: `Devices=N` on the operation,
`n_devices` in the parameter struct): the one-shot call shards the volume over devices 0..N-1 with peer copies between the
engines; with pyramid_levels > 1 the coarse levels run on device 0 and the finest level on the slab group.  Every result
must equal the one-device result bit for bit.

    python tests/multi_gpu/run_inprocess_devices.py [--devices 8]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=8)
    a = ap.parse_args()
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, demons_volume
    n = a.devices
    shape, sp = (16 * n, 96, 96), (1.0, 1.0, 1.0)
    f, m = phantom.make_pair(shape, spacing=sp, amplitude_mm=4.0)
    F, M = f.numpy(), m.numpy()
    ok = True
    cases = [("plain fp64_strict", dict(field_mode="fp64_strict")),
             ("symmetric mixed", dict(field_mode="mixed", use_symmetric_force=True)),
             ("pyramid3 symmetric fp64_strict", dict(field_mode="fp64_strict", pyramid_levels=3, pyramid_iterations=(10, 10, 0),
                                                     use_symmetric_force=True))]
    for name, kw in cases:
        base = dict(max_iterations=8, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                    update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, device=0)
        one, h1, s1 = AlignViaDemons(AlignViaDemonsParams(**base, **kw), demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
        many, hn, sn = AlignViaDemons(AlignViaDemonsParams(**base, n_devices=n, **kw), demons_volume(M, *sp), demons_volume(F, *sp),
                                      return_stats=True)
        same = one is not None and many is not None and np.array_equal(one.data, many.data)
        rel = float(np.max(np.abs(h1 - hn) / h1))
        print(f"[{n} devices, one process] {name}: iterations {sn.iterations_done}/{s1.iterations_done} bit_identical={same} "
              f"mse_rel={rel:.2e} loop_ms {s1.loop_ms:.1f} -> {sn.loop_ms:.1f}", flush=True)
        ok = ok and same and rel < 1e-6 and sn.iterations_done == s1.iterations_done
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
