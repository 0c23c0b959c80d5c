"""run_sp vs run_sp_resident (result left in HBM) on the c2 band-pass job and the c4 Monte-Carlo batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import heavi_b200 as hv
import circuits

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

m, f = circuits.build_c2(hv, "bandpass", nF=100_001)
a = t(lambda: m.run_sp(f)); b = t(lambda: m.run_sp_resident(f))
print("c2 band-pass 100 001 points: run_sp %.0f us (%.2e/s), run_sp_resident %.0f us (%.2e/s)" % (a * 1e6, 100001 / a, b * 1e6, 100001 / b))
m4, f4, mc = circuits.build_c4(hv)
np.random.seed(7); vals = mc.draw(1250)
a = t(lambda: m4.run_sp_batch(f4, mc._random_numbers, vals), 5); b = t(lambda: m4.run_sp_resident(f4, mc._random_numbers, vals), 5)
n = 1250 * len(f4)
print("c4 1250 x 2001 points: run_sp_batch %.2f ms (%.2e/s), run_sp_resident %.2f ms (%.2e/s)" % (a * 1e3, n / a, b * 1e3, n / b))
