"""TEST infrastructure (not product code): torch.distributed restatement of how registrations and z-slabs are spread over
ranks, used by tests/test_parallel_gloo.py to emulate the schedule of csrc/slab.cu on CPU with gloo.  The product data
plane is csrc/slab.cu (IPC peer stores + flag counters); nothing under dicomautomaton_b200/ imports this file.

The Demons path shards in two ways:
  * independent registrations (BASELINE config 5): round-robin over ranks, NO data-path collective -- `shard_cases`;
  * one large volume cut into contiguous z-slabs with a neighbour exchange per stencil stage (config 4) --
    `slab_partition`, `exchange_z_halos`; the MSE needs one all-reduce of (sum of squared differences, count).
Everything here works on CPU tensors with the gloo backend (tests) and on CUDA tensors with NCCL (bench).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_cases(n_cases: int, world: int, rank: int) -> List[int]:
    """Round-robin assignment of independent registrations to ranks (config 5: '8 per GPU round-robin')."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_cases, world))


def slab_partition(slices: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced z-slabs [z0, z1) per rank; earlier ranks take the remainder. Empty slabs are allowed only
    when world > slices."""
    base, rem = divmod(slices, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((z, z + n))
        z += n
    return out


def smoothing_radius(sigma_mm: float, pxl: float) -> int:
    """max(1, int(3 sigma / pxl)): reference src/Alignment_Demons.cc:574-576."""
    return max(1, int(3.0 * sigma_mm / pxl)) if sigma_mm > 0.0 else 0


def exchange_z_halos(slab: torch.Tensor, halo: int, rank: int, world: int, group=None) -> torch.Tensor:
    """slab: (..., nz_local, rows, cols) owned planes of this rank. Returns the slab extended by `halo` planes on each
    side that has a neighbour (global faces get none): lower neighbour's last planes first, then own, then the upper
    neighbour's first planes.  Point-to-point only (batch_isend_irecv): the z-neighbour exchange of SURVEY 8(e)."""
    if halo <= 0 or world == 1:
        return slab
    nz = slab.shape[-3]
    if nz < halo:
        raise ValueError("slab thinner than the halo: use fewer ranks")
    ops, lo, hi = [], None, None
    if rank > 0:
        lo = torch.empty_like(slab[..., :halo, :, :]).contiguous()
        ops.append(dist.P2POp(dist.isend, slab[..., :halo, :, :].contiguous(), rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, lo, rank - 1, group))
    if rank < world - 1:
        hi = torch.empty_like(slab[..., nz - halo:, :, :]).contiguous()
        ops.append(dist.P2POp(dist.isend, slab[..., nz - halo:, :, :].contiguous(), rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, hi, rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = [p for p in (lo, slab, hi) if p is not None]
    return torch.cat(parts, dim=-3)


def allreduce_mse(sum_sq: float, count: int, device="cpu", group=None) -> Tuple[float, int]:
    """Global MSE terms: one all-reduce of (sum of squared differences, valid count) (cc:334-342 across slabs)."""
    t = torch.tensor([float(sum_sq), float(count)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t[0]), int(round(float(t[1])))


def reduce_timing(ms_local: float, units_local: float, device="cpu", group=None) -> Tuple[float, float]:
    """(max over ranks of the elapsed time, sum over ranks of the units processed): the bench contract."""
    if not (dist.is_available() and dist.is_initialized()):
        return ms_local, units_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t[0]), float(u[0])
