"""e2e timing of the c5 ladder through Model.run_sp for a few host-pipeline settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import heavi_b200 as hv
import circuits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m, _ = circuits.build_c5(hv, nF=8)
f = np.linspace(0.5e9, 6e9, n)

def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3

print("run_sp            %.2f ms" % timeit(lambda: float(m.run_sp(f)._S[0, 0, -1].real)))
cn = m._compiled()
print("cn.run            %.2f ms" % timeit(lambda: cn.run(f)))
out = cn.run(f)
print("cn.run(out=)      %.2f ms" % timeit(lambda: cn.run(f, out=out)))
consts, prefs, tables, samples = cn.resolve(f)
print("run_host(out=)    %.2f ms" % timeit(lambda: cn.handle.run_host(f, consts, prefs, tables, samples, out)))
d = torch.empty(out.size, dtype=torch.complex128, device="cuda")
h = torch.from_numpy(out.reshape(-1))
def cp():
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("plain D2H %d MB  %.2f ms (pinned=%s)" % (out.nbytes >> 20, timeit(cp), h.is_pinned()))
print("f.astype(f32)     %.2f ms" % timeit(lambda: f.astype(np.float32)))
