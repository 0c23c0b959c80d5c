"""README demo (config 1, has an SMD part) at 1 M points: closed form on the device vs host-tabulated callable."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
from heavi_b200.params import SMDImpedance
pts = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
for tabulate in (False, True):
    m, _ = circuits.build_c1(hv)
    if tabulate:                                   # what every callable parameter costs: evaluated by numpy per run
        for c in m.components:
            for v in vars(c).values():
                if isinstance(v, SMDImpedance):
                    v._inf = 1e300
    f = np.linspace(1.8e9, 2.2e9, pts)
    for _ in range(2):
        m.run_sp(f)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); m.run_sp(f); ts.append(time.perf_counter() - t0)
    print("c1, %d points, SMD part %s: %.2f ms -> %.3e solves/s" % (pts, "tabulated on the host" if tabulate else "closed form on the device", np.median(ts) * 1e3, pts / np.median(ts)))
