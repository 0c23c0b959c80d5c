"""Ladders of 1500 / 2000 sections on the banded kernel: parity against the CPU oracle and a 200 k-point sweep."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, time
import heavi_b200 as hv, circuits
from heavi_b200.engine import CompiledNetwork
from oracle import mna_oracle as orc
for K,reach in ((1500,2),(2000,1)):
    m,f = circuits.build_band_ladder(hv,K=K,reach=reach,nports=3,seed=9,nF=5)
    t0=time.time(); cn=CompiledNetwork(m); t1=time.time()
    ki=cn.handle.kernel_info()
    S=cn.run(f)[0]
    So=orc.run_sp(m.records(), m.n_nodes, len(m.sources), f, method="lu_batched")
    print("K=%d reach=%d n=%d band=%d class=%d compile %.2fs  max|dS|=%.2e"%(K,reach,cn.plan.n,cn.plan.band,ki["kernel_class"],t1-t0,np.max(np.abs(S-So))))
    fb=np.linspace(0.5e9,6e9,200000)
    for _ in range(2): cn.run(fb)
    t0=time.perf_counter(); cn.run(fb); dt=time.perf_counter()-t0
    print("   200k points e2e %.1f ms -> %.3e solves/s"%(dt*1e3, 2e5/dt))
