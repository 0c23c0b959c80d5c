"""Library bar: torch.linalg.solve (cuSOLVER / cuBLAS batched LU, complex128) on
the SAME assembled systems, solve only (no assembly, no S epilogue), against our fused kernels
(assembly + solve + S).  Run on a GPU box:  python tools/cublas_compare.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import heavi_b200 as hv, circuits
from heavi_b200 import plan as pl
from heavi_b200.engine import DevicePlanHandle

def bench(name, m, f, reps=5):
    cn = m._compiled()
    table = cn.table
    p = pl.compile_plan(table, cn.plan.mode)                       # unpadded system of the same mode
    h = DevicePlanHandle(p)
    consts, prefs, tables, samples = pl.resolve_run(p, f)
    W = h.assemble_host(f, consts, prefs, tables, samples)[0]      # [nF, n, n+P]
    A = torch.from_numpy(np.ascontiguousarray(W[:, :, :p.n])).cuda()
    B = torch.from_numpy(np.ascontiguousarray(W[:, :, p.n:])).cuda()
    for _ in range(2): X = torch.linalg.solve(A, B)
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); X = torch.linalg.solve(A, B); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    t_lib = min(a.elapsed_time(b) for a, b in ev)
    # ours, device resident
    d_f = torch.from_numpy(f).cuda(); P = cn.plan.P
    c2, p2, t2, s2 = cn.resolve(f)
    d_S = torch.empty((1, P, P, len(f)), dtype=torch.complex128, device="cuda")
    d_smp = torch.from_numpy(s2).cuda(); d_st = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_tab = torch.from_numpy(t2).cuda() if t2.shape[0] else None
    for _ in range(2): cn.handle.run_device(d_f, c2, p2, d_tab, d_smp, d_S, d_st)
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cn.handle.run_device(d_f, c2, p2, d_tab, d_smp, d_S, d_st); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    t_ours = min(a.elapsed_time(b) for a, b in ev)
    print("%-22s n=%3d P=%d nF=%7d | torch.linalg.solve (solve only) %9.3f ms = %.3e solves/s | ours (assembly+solve+S) %9.3f ms = %.3e solves/s | x%.1f"
          % (name, p.n, P, len(f), t_lib, len(f) / t_lib * 1e3, t_ours, len(f) / t_ours * 1e3, t_lib / t_ours), flush=True)

if __name__ == "__main__":
    m, f = circuits.build_c2(hv, "bandpass"); bench("c2 band-pass", m, f)
    m, f = circuits.build_c2(hv, "lowpass"); bench("c2 low-pass", m, f)
    m, f = circuits.build_c3(hv, nF=100_000); bench("c3 touchstone", m, f)
    m, f = circuits.build_c5(hv, nF=2048); bench("c5 ladder", m, f)
