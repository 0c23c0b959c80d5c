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

import os
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


# --------------------------------------------------------------------------------------------------
# One result array for all ranks of a node: a POSIX shared-memory file every rank maps and page-locks, so
# that each GPU copies its block over its own PCIe link straight to its place in the final S[nS,P,P,nF]
# ("per-device D2H into one pinned array"; here across processes).  No inter-GPU
# transfer, no second copy on the host.  Buffers are pooled per shape because page-locking ~1 GB costs
# more than the sweep that fills it.
# --------------------------------------------------------------------------------------------------
_SHARED_POOL: dict = {}


class SharedResult:
    def __init__(self, name: str, shape, create: bool):
        import mmap
        self.name, self.shape = name, tuple(int(x) for x in shape)
        self.nbytes = int(np.prod(self.shape)) * 16
        path = "/dev/shm/" + name
        fd = os.open(path, os.O_RDWR | (os.O_CREAT if create else 0), 0o600)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self._mm = mmap.mmap(fd, self.nbytes)
        finally:
            os.close(fd)
        self.array = np.frombuffer(self._mm, dtype=np.complex128).reshape(self.shape)
        self.ptr = self.array.__array_interface__["data"][0]
        self.pinned = False

    def pin(self):
        """Page-lock the mapping for this process' CUDA context (asynchronous, full-rate D2H)."""
        if self.pinned:
            return
        try:
            import torch
            if not torch.cuda.is_available():
                return
            self.pinned = int(torch.cuda.cudart().cudaHostRegister(self.ptr, self.nbytes, 0)) == 0
        except Exception:
            self.pinned = False                          # pageable still works (staged copies)

    def unlink(self):
        try:
            os.unlink("/dev/shm/" + self.name)
        except OSError:
            pass


def _shared_result(shape, group, dst: int, owner=None) -> SharedResult:
    """Every rank's view of one pooled shared array of `shape` (created by rank `dst` on first use).  The pool
    is keyed by shape AND owner (the compiled netlist): two models never share a buffer, repeated runs of one
    model reuse -- and overwrite -- theirs."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    key = (tuple(int(x) for x in shape), id(group), dst, id(owner))
    if key in _SHARED_POOL:
        ref, res = _SHARED_POOL[key]
        if owner is None or ref() is owner:               # id() of a dead object may be reused
            return res
        del _SHARED_POOL[key]
    box = [None]
    if rank == dst:
        box[0] = "hvb_%d_%d_%s" % (os.getpid(), len(_SHARED_POOL), "x".join(str(int(x)) for x in shape))
        res = SharedResult(box[0], shape, create=True)
    dist.broadcast_object_list(box, src=dst, group=group)
    if rank != dst:
        res = SharedResult(box[0], shape, create=False)
    res.pin()
    dist.barrier(group)
    if rank == dst:
        res.unlink()                                     # the mapping lives on; the name does not
    import weakref
    try:
        ref = weakref.ref(owner) if owner is not None else (lambda: None)
    except TypeError:
        ref = (lambda o=owner: o)
    _SHARED_POOL[key] = (ref, res)
    return res


def run_sp_distributed(net, frequencies, drivers=None, values=None, group=None, dst: int = 0, gather: str = "shared",
                       writer=None):
    """`run_sp` across the ranks of an initialised process group (one GPU per rank).  Rank `dst`
    returns S[nS,P,P,nF]; the others return None.

    gather="shared" (ranks on one node, the default): every rank's GPU writes its block directly into one
    page-locked shared-memory array; the array rank `dst` gets back is pooled per (model, shape) and is
    overwritten by the next call of the same model with the same shape (copy it to keep it).  gather="collective": `dist.gather` of the blocks (NCCL or gloo), then
    one assembly on `dst` -- for ranks that do not share a host.

    `writer(shard, f_block, values_block, result)` (tests): fills this rank's block of `result.array` on the
    host instead of the GPU call; `net` may then be any object with a `P` attribute."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    f = np.ascontiguousarray(np.asarray(frequencies, dtype=np.float64))
    compiled = net._compiled() if writer is None else None
    nS = 1 if values is None else int(np.asarray(values).shape[0])
    nF, P = f.shape[0], (compiled.plan.P if writer is None else int(net.P))
    total = nS * nF                                      # every rank picks the kernel variant from the WHOLE job

    if gather == "collective":
        def compute(fb, vb):
            return compiled.run(fb, drivers=drivers, values=vb, batch_points=total)
        sh, blk = run_sharded(compute, f, values, world, rank)
        return gather_to_root(blk, sh, nS, nF, group=group, dst=dst)
    if gather != "shared":
        raise ValueError(gather)
    res = _shared_result((nS, P, P, nF), group, dst, owner=compiled if compiled is not None else net)
    sh = plan_shard(nS, nF, world, rank)
    if sh.size > 0 and writer is not None:
        if sh.axis == "sample":
            writer(sh, f, None if values is None else np.asarray(values)[sh.lo:sh.hi], res)
        else:
            writer(sh, f[sh.lo:sh.hi], values, res)
    elif sh.size > 0:
        if sh.axis == "sample":
            vb = None if values is None else np.asarray(values)[sh.lo:sh.hi]
            compiled.run(f, drivers=drivers, values=vb, batch_points=total,
                         out_ptr=res.ptr + sh.lo * P * P * nF * 16, out_ld=0)
        else:
            compiled.run(f[sh.lo:sh.hi], drivers=drivers, values=values, batch_points=total,
                         out_ptr=res.ptr + sh.lo * 16, out_ld=nF)
    dist.barrier(group)                                  # every block has landed (each run() returns after its copies)
    return res.array if rank == dst else None


def merge_moments(blocks, counts):
    """Merge per-block Monte-Carlo statistics [6,P,P,nF] (mean, population std of |S|, |S|^2, sqrt(|S|^2) --
    `hvb_run_stats_host`, sparam.py:215-243) of sample blocks with `counts` samples each into the statistics
    of all samples: pooled mean, centred second moments (M2 = sum n_b (std_b^2 + (mean_b - mean)^2))."""
    counts = np.asarray(counts, dtype=np.float64)
    N = counts.sum()
    blocks = [np.asarray(b, dtype=np.float64) for b in blocks]
    out = np.empty_like(blocks[0])
    for q in range(3):
        means = np.stack([b[2 * q] for b in blocks])
        stds = np.stack([b[2 * q + 1] for b in blocks])
        w = counts.reshape((-1,) + (1,) * (means.ndim - 1))
        mean = (w * means).sum(axis=0) / N
        m2 = (w * (stds ** 2 + (means - mean) ** 2)).sum(axis=0)
        out[2 * q] = mean
        out[2 * q + 1] = np.sqrt(m2 / N)
    return out


def run_mc_stats_distributed(net, frequencies, drivers, values, group=None, dst: int = 0, stats=None):
    """Monte-Carlo statistics over the ranks of a process group: samples are sharded, each rank reduces its
    block on its GPU (`run_stats`), and the [6,P,P,nF] moments are combined with two small sum-reductions
    (NCCL over NVLink, or gloo) -- the one place this path has a collective (SURVEY.md section 8e/f1;
    reference: numeric.py:498-507, sparam.py:215-295).  Rank `dst` returns the moments, the others None.
    `stats(values_block) -> [6,P,P,nF]` (tests) replaces the GPU call; `net` may then be None."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    f = np.ascontiguousarray(np.asarray(frequencies, dtype=np.float64))
    values = np.asarray(values, dtype=np.float64)
    nS = values.shape[0]
    lo, hi = split_range(nS, world, rank)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if stats is None:
        compiled = net._compiled()
        mom = compiled.run_stats(f, list(drivers), values[lo:hi]) if hi > lo else np.zeros((6, compiled.plan.P, compiled.plan.P, f.shape[0]))
    else:
        mom = np.asarray(stats(values[lo:hi]), dtype=np.float64)
    P, nF = mom.shape[1], mom.shape[3]
    n = float(hi - lo)
    means = torch.from_numpy(np.ascontiguousarray(mom[0::2]) * n).to(dev)            # sum of n_b * mean_b
    dist.all_reduce(means, op=dist.ReduceOp.SUM, group=group)
    mean = means.cpu().numpy() / nS
    m2 = torch.from_numpy(n * (mom[1::2] ** 2 + (mom[0::2] - mean) ** 2)).to(dev)    # centred on the pooled mean
    dist.reduce(m2, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    out = np.empty((6, P, P, nF))
    out[0::2] = mean
    out[1::2] = np.sqrt(m2.cpu().numpy() / nS)
    return out
