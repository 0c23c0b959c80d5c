"""End-to-end Model.run_sp on a mid-size (n = K+1) narrow-banded ladder: register kernel vs banded variant."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
K = int(sys.argv[1]) if len(sys.argv) > 1 else 19
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
m, _ = circuits.build_band_ladder(hv, K=K, reach=1, nports=2, seed=21)
f = np.linspace(0.5e9, 6e9, pts)
for _ in range(3):
    m.run_sp(f)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); S = m.run_sp(f); ts.append(time.perf_counter() - t0)
cn = m._compiled()
print("n=%d points=%d NO_BAND=%s: %.2f ms -> %.3e solves/s (banded variant %s)" % (
    cn.plan.n_real, pts, os.environ.get("HVB_NO_BAND"), np.median(ts) * 1e3, pts / np.median(ts), "used" if cn._band else "not used"))
