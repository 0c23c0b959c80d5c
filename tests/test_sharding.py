"""CPU: the multi-GPU path's host logic -- partitioning and the world_size-2 gather over gloo.
The per-rank compute is injected (here: the numpy oracle), so no GPU is needed."""
import os
import socket

import numpy as np
import pytest

from heavi_b200 import sharding as sh


def test_split_range_covers_everything():
    for total in (0, 1, 7, 100_001, 2001):
        for world in (1, 2, 3, 4, 8):
            edges = [sh.split_range(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_plan_shard_axis_choice():
    assert sh.plan_shard(10_000, 2001, 8, 3).axis == "sample"
    assert sh.plan_shard(1, 1_000_000, 8, 3) == sh.Shard("freq", 375_000, 500_000)
    with pytest.raises(ValueError):
        sh.plan_shard(1, 10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    import heavi_b200 as hv
    import circuits
    from oracle import mna_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode == "freq":
            m, f = circuits.build_c2(hv, "bandpass", nF=301)
            values, drivers = None, None
        else:
            m, f, mc = circuits.build_c4(hv, nF=41)
            np.random.seed(3)
            values, drivers = mc.draw(5), mc._random_numbers

        def compute(fb, vb):
            if vb is None:
                return orc.run_sp(m.records(), m.n_nodes, len(m.sources), fb, method="lu_batched")[None]
            out = []
            for row in vb:
                for d, v in zip(drivers, row):
                    d._value = v
                out.append(orc.run_sp(m.records(), m.n_nodes, len(m.sources), fb, method="lu_batched"))
            return np.stack(out)

        nS = 1 if values is None else values.shape[0]
        shard, blk = sh.run_sharded(compute, f, values, world, rank)
        full = sh.gather_to_root(blk, shard, nS, len(f))
        if rank == 0:
            whole = compute(f, values)
            q.put(bool(np.array_equal(full, whole)))          # bit-identical to the unsharded run
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["freq", "sample"])
def test_two_rank_gloo_gather_is_bit_identical(mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _moments(S):
    a = np.abs(S)
    pw = a ** 2
    return np.stack([a.mean(axis=0), a.std(axis=0), pw.mean(axis=0), pw.std(axis=0), np.sqrt(pw).mean(axis=0), np.sqrt(pw).std(axis=0)])


def test_merge_moments_equals_statistics_of_all_samples():
    rng = np.random.default_rng(5)
    S = rng.normal(size=(37, 2, 2, 11)) + 1j * rng.normal(size=(37, 2, 2, 11))
    whole = _moments(S)
    for cuts in ([0, 37], [0, 5, 37], [0, 1, 2, 20, 36, 37], [0, 12, 12, 37]):
        blocks = [(S[a:b], b - a) for a, b in zip(cuts, cuts[1:]) if b > a]
        merged = sh.merge_moments([_moments(b) for b, _ in blocks], [n for _, n in blocks])
        assert np.max(np.abs(merged - whole)) < 1e-13


def _worker_shared(rank, world, port, mode, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import types
    import torch.distributed as dist
    import heavi_b200 as hv
    import circuits
    from oracle import mna_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode == "freq":
            m, f = circuits.build_c2(hv, "bandpass", nF=301)
            values, drivers = None, None
        else:
            m, f, mc = circuits.build_c4(hv, nF=41)
            np.random.seed(3)
            values, drivers = mc.draw(5), mc._random_numbers

        def compute(fb, vb):
            if vb is None:
                return orc.run_sp(m.records(), m.n_nodes, len(m.sources), fb, method="lu_batched")[None]
            out = []
            for row in vb:
                for d, v in zip(drivers, row):
                    d._value = v
                out.append(orc.run_sp(m.records(), m.n_nodes, len(m.sources), fb, method="lu_batched"))
            return np.stack(out)

        def writer(shard, fb, vb, res):                       # what the GPU does with out_ptr / out_ld, on the host
            blk = compute(fb, vb)
            if shard.axis == "sample":
                res.array[shard.lo:shard.hi] = blk
            else:
                res.array[..., shard.lo:shard.hi] = blk

        ok = True
        for _ in range(2):                                    # second call reuses the pooled shared array
            full = sh.run_sp_distributed(types.SimpleNamespace(P=2), f, drivers, values, writer=writer)
            if rank == 0:
                ok = ok and bool(np.array_equal(full, compute(f, values)))
        if mode == "sample":
            mom = sh.run_mc_stats_distributed(None, f, drivers, values, stats=lambda vb: _moments(compute(f, vb)))
            if rank == 0:
                ok = ok and float(np.max(np.abs(mom - _moments(compute(f, values))))) < 1e-13
        if rank == 0:
            q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["freq", "sample"])
def test_two_rank_shared_memory_result_and_moment_reduction(mode):
    """Every rank writes its block into ONE shared-memory array (what each GPU does over its own PCIe link);
    Monte-Carlo moments are combined with two sum-reductions."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_shared, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
