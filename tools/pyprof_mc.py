"""Where the time goes in Network.run_mc_stats / run_mc on config 4 (GPU box)."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
m, f, mc = circuits.build_c4(hv)
N = 1250
def t(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3
print("run_mc_stats %.3f ms" % t(lambda: m.run_mc_stats(f, mc, N)))
print("run_mc       %.3f ms" % t(lambda: m.run_mc(f, mc, N)))
print("mc.draw      %.3f ms" % t(lambda: mc.draw(N)))
cn = m._compiled()
vals = mc.draw(N)
print("cn.run_stats %.3f ms" % t(lambda: cn.run_stats(f, mc._random_numbers, vals)))
print("cn.run       %.3f ms" % t(lambda: cn.run(f, mc._random_numbers, vals)))
print("resolve      %.3f ms" % t(lambda: cn.resolve(f, mc._random_numbers, vals)))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): m.run_mc_stats(f, mc, N)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(10)
