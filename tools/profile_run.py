#!/usr/bin/env python3
"""Small driver for ncu captures: one registration of `--iters` iterations on the bench workload (inputs made on the
GPU, engine-level API, no host copies).  Not a benchmark: numbers printed under a profiler are never reported."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="fp64")
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--shape", default="256,512,512")
    ap.add_argument("--standard", action="store_true", help="non-diffeomorphic Demons")
    ap.add_argument("--unfused", action="store_true")
    a = ap.parse_args()
    import torch
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import AlignViaDemonsParams, DemonsEngine
    shape = tuple(int(v) for v in a.shape.split(","))
    f, m = phantom.make_pair(shape, device="cuda")
    par = AlignViaDemonsParams(max_iterations=a.iters, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                               update_field_smoothing_sigma=1.5, use_diffeomorphic=not a.standard, verbosity=0,
                               field_mode=a.mode, fused=not a.unfused)
    eng = DemonsEngine.empty(shape, (1.0, 1.0, 1.0), par)
    eng.load_device(f.data_ptr(), m.data_ptr())
    hist, st = eng.run(a.iters)
    print(f"mode={a.mode} iters={st.iterations_done} launches={st.kernel_launches} loop_ms={st.loop_ms:.3f} "
          f"mse={hist[-1]:.6f}")
    eng.close()


if __name__ == "__main__":
    main()
