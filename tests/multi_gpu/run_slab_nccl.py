#!/usr/bin/env python3
"""Multi-process z-slab check (one rank per GPU, NCCL halos), to be launched on a multi-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu/run_slab_nccl.py [--shape 64,96,128] [--iters 10]

Every rank also runs the whole-volume engine on its own GPU and compares its slab of the sharded result with it:
the fields must be bit-identical.  Prints one line per rank and exits non-zero on mismatch."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="64,96,128")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--mode", default="fp64_strict")
    ap.add_argument("--symmetric", action="store_true", help="use_symmetric_force (extension a11): W needs a z halo")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, DemonsSlabGroup, demons_volume

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # plumbing only: the 128-byte NCCL id; the data path uses the library's own NCCL
    shape = tuple(int(v) for v in a.shape.split(","))
    sp = (1.0, 1.0, 1.0)
    f, m = phantom.make_pair(shape, spacing=sp)
    F, M = f.numpy(), m.numpy()
    ok = True
    for diffeo in (False, True):
        par = AlignViaDemonsParams(max_iterations=a.iters, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5,
                                   update_field_smoothing_sigma=1.5, use_diffeomorphic=diffeo, verbosity=0,
                                   field_mode=a.mode, device=local, use_symmetric_force=a.symmetric)
        ident = [DemonsSlabGroup.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        grp = DemonsSlabGroup.nccl(shape, sp, par, ident[0], rank, world, local)
        grp.load(F, M)
        hist, st = grp.run()
        z0, z1 = grp.owned_range()
        Ds = grp.get_deformation()
        hb = grp.halo_bytes
        xs = grp.exchange_stats()
        grp.close()
        whole, hw, _ = AlignViaDemons(par, demons_volume(M, *sp), demons_volume(F, *sp), return_stats=True)
        same = np.array_equal(whole.data[z0:z1], Ds)
        mse_rel = float(np.max(np.abs(hist - hw) / hw))
        print(f"[rank {rank}/{world}] diffeo={diffeo} slices [{z0},{z1}) iters={st.iterations_done} loop_ms={st.loop_ms:.2f} "
              f"halo_bytes={hb} bit_identical={same} mse_rel={mse_rel:.2e} overlapped={xs['overlapped']} "
              f"exchange_ms={xs['exchange_ms']:.2f} transport='{xs['transport']}'", flush=True)
        ok = ok and same and mse_rel < (1e-6 if a.mode in ("fp32", "mixed") else 1e-12) and st.iterations_done == a.iters
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
