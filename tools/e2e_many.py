"""Serial `run_sp` calls vs `run_sp_many` on the configs[1] filter set (end to end, host buffers)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import bench

jobs = [(j[0], j[1]) for j in bench.make_jobs(hv, "c2")]
pts = sum(len(f) for _, f in jobs)
for name, fn in (("serial run_sp", lambda: [m.run_sp(f) for m, f in jobs]), ("run_sp_many", lambda: hv.run_sp_many(jobs))):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    print("%-14s median %.3f ms  min %.3f ms  max %.3f ms  -> %.3e solves/s" % (name, np.median(ts) * 1e3, min(ts) * 1e3, max(ts) * 1e3, pts / np.median(ts)), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    hv.run_sp_many(jobs)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
