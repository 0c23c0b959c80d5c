"""Multi-GPU sharding of a sweep: one process per GPU (torchrun), no data-path collective.

Every (sample, frequency) point is independent, so the only multi-GPU logic
is index arithmetic plus an optional gather of the results:

* Monte-Carlo / sweep batches (nS >= world): rank r takes a contiguous block of samples and
  produces a contiguous [nS_r, P, P, nF] block.
* plain sweeps: rank r takes a contiguous frequency slab and produces P*P rows of nF_r values.

Results are bit-identical for any world size because a point's arithmetic never depends on which
rank or chunk it lands in.  `gather_to_root` assembles the full array on rank 0 either through host
memory (gloo, or NCCL-free per-device D2H + a host gather) or, when the process group is NCCL,
through `dist.gather` on the device followed by one D2H.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    axis: str          # "sample" or "freq"
    lo: int
    hi: int

    @property
    def size(self) -> int:
        return self.hi - self.lo


def split_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous near-equal split: the first `total % world` ranks get one extra element."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def plan_shard(nS: int, nF: int, world: int, rank: int) -> Shard:
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {world}/{rank}")
    if nS >= world:
        return Shard("sample", *split_range(nS, world, rank))
    return Shard("freq", *split_range(nF, world, rank))


def run_sharded(compute, f: np.ndarray, values: np.ndarray | None, world: int, rank: int):
    """Run this rank's share.  `compute(f_block, values_block) -> S[nS_b, P, P, nF_b]` is the
    single-GPU entry (`CompiledNetwork.run`).  Returns (Shard, S_block)."""
    f = np.asarray(f, dtype=np.float64)
    nS = 1 if values is None else int(np.asarray(values).shape[0])
    sh = plan_shard(nS, f.shape[0], world, rank)
    if sh.axis == "sample":
        vb = None if values is None else np.asarray(values)[sh.lo:sh.hi]
        return sh, compute(f, vb)
    return sh, compute(f[sh.lo:sh.hi], values)


def gather_to_root(S_block: np.ndarray, shard: Shard, nS: int, nF: int, group=None, dst: int = 0):
    """Assemble S[nS,P,P,nF] on rank `dst` from every rank's block; other ranks get None."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    P = S_block.shape[1]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    # blocks differ in size by at most one row along the sharded axis: pad to the largest
    sizes = [plan_shard(nS, nF, world, r).size for r in range(world)]
    pad = max(sizes)
    if shard.axis == "sample":
        buf = np.zeros((pad, P, P, nF), dtype=np.complex128)
        buf[:shard.size] = S_block
    else:
        buf = np.zeros((nS, P, P, pad), dtype=np.complex128)
        buf[..., :shard.size] = S_block
    t = torch.view_as_real(torch.from_numpy(buf)).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, outs, dst=dst, group=group)
    if rank != dst:
        return None
    full = np.empty((nS, P, P, nF), dtype=np.complex128)
    for r, o in enumerate(outs):
        sh = plan_shard(nS, nF, world, r)
        blk = torch.view_as_complex(o.cpu().contiguous()).numpy()
        if sh.axis == "sample":
            full[sh.lo:sh.hi] = blk[:sh.size]
        else:
            full[..., sh.lo:sh.hi] = blk[..., :sh.size]
    return full


def run_sp_distributed(net, frequencies, drivers=None, values=None, group=None, dst: int = 0):
    """`run_sp` across the ranks of an initialised process group (one GPU per rank).  Rank `dst`
    returns S[nS,P,P,nF]; the others return None."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    f = np.ascontiguousarray(np.asarray(frequencies, dtype=np.float64))
    compiled = net._compiled()
    nS = 1 if values is None else int(np.asarray(values).shape[0])

    def compute(fb, vb):
        # every rank picks the kernel variant from the size of the WHOLE job: bit-identical for any world size
        return compiled.run(fb, drivers=drivers, values=vb, batch_points=nS * f.shape[0])

    sh, blk = run_sharded(compute, f, values, world, rank)
    return gather_to_root(blk, sh, nS, f.shape[0], group=group, dst=dst)
