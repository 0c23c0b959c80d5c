import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
m, f = circuits.build_c2(hv, "bandpass", nF=100_001)
for _ in range(30):
    t0 = time.perf_counter(); S = m.run_sp(f); t1 = time.perf_counter()
    sys.stderr.write("run_sp %.1f us\n" % ((t1 - t0) * 1e6))
