"""Tiny driver for ncu: a few device-resident launches of one workload's dominant kernel.
    python tools/prof_run.py c2|c1|c3|c4|c5 [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import heavi_b200 as hv
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
jobs = bench.make_jobs(hv, wl)
m, f, drivers, values = jobs[0]
if os.environ.get("PROF_POINTS"):                 # a rank's share of a strong-scaled sweep
    f = f[:int(os.environ["PROF_POINTS"])]
if wl == "c4":
    values = values[:1250]
cn = m._compiled()
consts, prefs, tables, samples = cn.resolve(f, drivers, values)
P = cn.plan.P
d_f = torch.from_numpy(f).to(dev)
d_tab = torch.from_numpy(tables).to(dev) if tables.shape[0] else None
d_smp = torch.from_numpy(samples).to(dev)
d_S = torch.empty((samples.shape[0], P, P, len(f)), dtype=torch.complex128, device=dev)
d_st = torch.zeros(1, dtype=torch.int32, device=dev)
ev = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); cn.handle.run_device(d_f, consts, prefs, d_tab, d_smp, d_S, d_st); b.record()
    ev.append((a, b))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
pts = len(f) * samples.shape[0]
print(wl, cn.handle.kernel_info(), "n=%d mode=%s" % (cn.plan.n, cn.plan.mode))
print("launch ms:", ["%.4f" % x for x in ms], "-> %.3e solves/s, %.2f TFLOP/s (LU-equivalent)" % (
    pts / (min(ms) * 1e-3), cn.plan.flops_per_solve() * pts / (min(ms) * 1e-3) / 1e12), "status", int(d_st.item()))
