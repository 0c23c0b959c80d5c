"""Where does the end-to-end time of run_sp go?  (run on a GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import heavi_b200 as hv, circuits
from heavi_b200 import engine

m, f = circuits.build_c2(hv, "bandpass")
cn = m._compiled()
for _ in range(3): m.run_sp(f)
def T(fn, n=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
print("run_sp total            %8.1f us" % T(lambda: m.run_sp(f)))
print("resolve_run             %8.1f us" % T(lambda: cn.resolve(f)))
print("pinned_empty            %8.1f us" % T(lambda: engine.pinned_empty((1, 2, 2, len(f)))))
consts, prefs, tables, samples = cn.resolve(f)
out = engine.pinned_empty((1, 2, 2, len(f)))
print("hvb_run_host (pinned)   %8.1f us" % T(lambda: cn.handle.run_host(f, consts, prefs, tables, samples, out)))
out2 = np.empty((1, 2, 2, len(f)), dtype=np.complex128)
print("hvb_run_host (pageable) %8.1f us" % T(lambda: cn.handle.run_host(f, consts, prefs, tables, samples, out2)))
print("f.astype(float32)       %8.1f us" % T(lambda: np.array(f).astype(np.float32)))
d = torch.empty((4, len(f)), dtype=torch.complex128, device="cuda")
h = torch.empty((4, len(f)), dtype=torch.complex128, pin_memory=True)
print("plain D2H 6.4MB pinned  %8.1f us" % T(lambda: h.copy_(d, non_blocking=True)))
for mb in ("4", "2", "1024"):
    os.environ["HVB_CHUNK_MB"] = mb
    print("hvb_run_host chunk %4s MB %8.1f us" % (mb, T(lambda: cn.handle.run_host(f, consts, prefs, tables, samples, out))))
