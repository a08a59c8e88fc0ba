#!/usr/bin/env python3
"""Full-length reference results as small fixtures, so the GPU suite can prove parity at BASELINE.json's real sizes and
iteration counts without spending CPU-oracle minutes on the GPU box.

    python tests/golden/make_full_size_fixtures.py [case ...]        (needs oracle/_ref, i.e. /root/reference)

For every case the UNMODIFIED reference engine (oracle/_ref/libdemons_ref.so, compiled from /root/reference by
oracle/build_ref.py) registers a bit-reproducible phantom pair (`phantom.make_pair_exact`: IEEE-exact operations only,
so the GPU box regenerates byte-identical inputs) and the following is stored in tests/golden/full/full_<case>.npz:

    sha256 of the fixed and moving volumes      -> the test first proves it has the same inputs
    sha256 of the reference's deformation field -> `fp64_strict` must hash to the same value (bit identity at full size)
    the field on a coarse sub-lattice, the MSE history, the iteration count

The other precision modes are then compared, voxel by voxel, with the strict GPU result that has just been proven
identical to the reference's.  tests/test_gpu_full_length.py consumes these files.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {
    # BASELINE config 1 exactly: 128^3, single-resolution standard Demons, sigma_d 1.0 mm, 100 iterations, threshold 0
    "config1_128": dict(shape=(128, 128, 128), sub=8, params=dict(
        max_iterations=100, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.0,
        update_field_smoothing_sigma=0.5, use_diffeomorphic=False)),
    # BASELINE config 2 parameters (diffeomorphic, sigma 1.5 mm, 200 iterations) at 1/8 of the volume
    "config2_256x256x128": dict(shape=(128, 256, 256), sub=8, params=dict(
        max_iterations=200, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
        update_field_smoothing_sigma=1.5, use_diffeomorphic=True)),
    # BASELINE config 2 exactly: 512x512x256, 200 iterations
    "config2_512x512x256": dict(shape=(256, 512, 512), sub=16, params=dict(
        max_iterations=200, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
        update_field_smoothing_sigma=1.5, use_diffeomorphic=True)),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).hexdigest()


def main():
    import oracle
    from dicomautomaton_b200 import phantom
    if not oracle.available("reference"):
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run oracle/build_ref.py")
    names = sys.argv[1:] or list(CASES)
    for name in names:
        c = CASES[name]
        f, m = phantom.make_pair_exact(c["shape"])
        F, M = f.numpy(), m.numpy()
        t = time.time()
        D, hist, n = oracle.demons_run(F, M, oracle.make_params(**c["params"]), kind="reference")
        dt = time.time() - t
        s = c["sub"]
        os.makedirs(os.path.join(HERE, "full"), exist_ok=True)
        out = os.path.join(HERE, "full", f"full_{name}.npz")
        np.savez_compressed(out, shape=np.array(c["shape"]), sub=s, params=np.array(sorted((k, str(v)) for k, v in c["params"].items())),
                            sha_fixed=sha(F), sha_moving=sha(M), sha_deformation=sha(D), deformation_sub=D[::s, ::s, ::s].copy(),
                            mse_history=hist[:n].copy(), iterations_done=n, max_abs=np.abs(D).max(axis=(0, 1, 2)),
                            generator="oracle/_ref (reference sycl_demons_engine), %d worker threads, %.1f s" % (oracle.worker_threads("reference"), dt))
        print(name, "->", out, f"{dt:.1f} s, iterations {n}, final MSE {hist[n - 1]:.6g}, max|D| {np.abs(D).max():.4f} mm", flush=True)


if __name__ == "__main__":
    main()
