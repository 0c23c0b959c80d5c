"""D2H rate into page-locked host memory with N ranks copying AT THE SAME TIME (torchrun, one rank per GPU):
what bounds the end-to-end rate when every GPU returns its block of S over its own PCIe link.
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe_ranks.py [MB]
Prints per-rank and aggregate GB/s for all ranks together, then for rank 0 alone."""
import os, sys, time
import torch, torch.distributed as dist

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 128.0
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(mb * 1e6)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)


def rate(reps=20):
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t) / 1e9


together = rate()
t = torch.tensor([together], dtype=torch.float64, device="cuda")
if world > 1:
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    rates = [float(x[0]) for x in allr]
else:
    rates = [together]
alone = None
for r in range(world):                      # one rank at a time
    if world > 1:
        dist.barrier()
    if r == rank and r == 0:
        alone = n * 20 / 1e9
        for _ in range(3):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        alone = alone / (time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
if rank == 0:
    print("pcie_probe_ranks: %d ranks, %.0f MB copies, D2H into pinned memory" % (world, mb))
    print("  all ranks at once: per rank GB/s = %s  aggregate = %.1f GB/s" % (" ".join("%.1f" % x for x in rates), sum(rates)))
    print("  rank 0 alone: %.1f GB/s" % alone)
if world > 1:
    dist.destroy_process_group()
