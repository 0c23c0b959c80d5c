"""cProfile of the Python side of Model.run_sp on the c2 band-pass job (GPU box)."""
import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
m, f = circuits.build_c2(hv, "bandpass", nF=100_001)
for _ in range(20):
    m.run_sp(f)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    m.run_sp(f)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
