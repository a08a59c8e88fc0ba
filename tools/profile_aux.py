#!/usr/bin/env python3
"""Driver for ncu captures of the kernels OUTSIDE the iteration loop (rows a13-a15, f1-f3, a12 of SURVEY section 8):
histogram matching, resampling / warping on another grid, pyramid restriction + prolongation, field text format.
One call of each at a clinically sized volume through the C ABI (host buffers).  Not a benchmark."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="128,256,256")
    a = ap.parse_args()
    from dicomautomaton_b200 import demons, field_io, phantom
    shape = tuple(int(v) for v in a.shape.split(","))
    f, m = phantom.make_pair(shape, device="cpu")
    f, m = f.numpy().reshape(shape), m.numpy().reshape(shape)
    rng = np.random.default_rng(5)
    D = (rng.standard_normal(shape + (3,)) * 1.5)
    t = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        r = fn()
        t[name] = round((time.perf_counter() - t0) * 1e3, 2)
        return r

    timed("histogram_match", lambda: demons.histogram_match(m, f))
    g_field = demons.make_grid(shape, (1.0, 1.0, 1.0))
    dshape = (shape[0] * 2 // 5, shape[1] * 2 // 5, shape[2] * 2 // 5)
    g_dose = demons.make_grid(dshape, (2.5, 2.5, 2.5), origin=(0.7, 0.4, 0.9))
    dose = rng.random(dshape, dtype=np.float32)
    timed("warp_image_grid", lambda: demons.warp_image_on_grid(g_field, D, g_dose, dose))
    timed("resample_to_grid", lambda: demons.resample_to_grid(g_dose, dose, g_field))
    vol = demons.demons_volume(f.reshape(shape + (1,)), 1.0, 1.0, 1.0)
    timed("warp_image", lambda: demons.warp_image_with_field(vol, D))
    coarse = timed("restrict_image", lambda: demons.restrict_image(vol))
    cf = rng.standard_normal(coarse.data.shape[:3] + (3,))
    timed("prolong_field", lambda: demons.prolong_field(cf, vol))
    vals = D.reshape(-1)[: 6 * 1024 * 1024]
    txt = timed("format_doubles", lambda: field_io.format_doubles(vals))
    back, used = timed("parse_doubles", lambda: field_io.parse_doubles(txt, vals.size))
    assert used == len(txt) - 1 and np.array_equal(back, vals)
    print("host wall ms per call (C ABI, pageable host buffers, first call each):", t)


if __name__ == "__main__":
    main()
