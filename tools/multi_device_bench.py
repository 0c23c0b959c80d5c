"""Single-process multi-device mode of the public API: `Model.run_sp(f)` on N visible GPUs without torchrun
(`hvb_run_host_multi`: one plan replica and host thread per device, every device copies its block straight
into the caller's one pinned array).  Prints solves/s for 1..N devices and checks that the results are
bit-identical.
    python tools/multi_device_bench.py [c5|c3|c4|c2] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import heavi_b200 as hv
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ndev = torch.cuda.device_count()
ref = None
for n in [d for d in (1, 2, 4, 8) if d <= ndev]:
    os.environ["HVB_DEVICES"] = str(n)
    jobs = bench.make_jobs(hv, wl)
    def once():
        out = []
        for (m, f, drivers, values) in jobs:
            out.append(m.run_sp(f)._S[None] if values is None else m.run_sp_batch(f, drivers, values))
        return out
    for _ in range(3):                    # warm-up: NVRTC compile, replica plans, pinned result buffers
        res = None
        res = once()
    t = time.perf_counter()
    for _ in range(reps):
        res = None                        # the previous result is released first: one pinned array in use at a time
        res = once()
    dt = (time.perf_counter() - t) / reps
    pts = sum(bench.job_points(j) for j in jobs)
    same = True if ref is None else all(np.array_equal(a, b) for a, b in zip(ref, res))
    if ref is None:
        ref = [r.copy() for r in res]
    print("%s devices=%d  %.2f ms per pass  %.3e solves/s  bit-identical to 1 device: %s" % (wl, n, dt * 1e3, pts / dt, same), flush=True)
