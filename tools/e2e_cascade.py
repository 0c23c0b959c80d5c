"""run_sp on a cascade of two-port S blocks (tests/circuits.build_twoport_cascade): banded variant vs register kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
sections = int(sys.argv[1]) if len(sys.argv) > 1 else 6
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
m, _ = circuits.build_twoport_cascade(hv, sections=sections)
f = np.linspace(0.5e9, 6e9, pts)
for _ in range(2):
    m.run_sp(f)
ts = []
for _ in range(4):
    t0 = time.perf_counter(); m.run_sp(f); ts.append(time.perf_counter() - t0)
cn = m._compiled()
ki = (cn._band[1] if cn._band else cn.handle).kernel_info()
print("%d sections, n=%d, %d points, NO_BAND=%s: %.2f ms -> %.3e solves/s (kernel class %d)" % (
    sections, cn.plan.n_real, pts, os.environ.get("HVB_NO_BAND"), np.median(ts) * 1e3, pts / np.median(ts), ki["kernel_class"]))
