"""Kernel-only timing of banded ladders of a given reach for the kernel-C variants (GPU box).
   python tools/band_bench.py REACH K NPORTS POINTS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import heavi_b200 as hv
import circuits
from heavi_b200.engine import CompiledNetwork
reach, K, nports, pts = (int(x) for x in sys.argv[1:5])
m, _ = circuits.build_band_ladder(hv, K=K, reach=reach, nports=nports, seed=3)
f = np.linspace(0.5e9, 6e9, pts)
cn = CompiledNetwork(m)
if os.environ.get("HVB_BAND_VARIANT") == "1" and cn._band is not False and cn._band_variant():
    cn.plan, cn.handle = cn._band_variant()          # time the banded variant of a 9..32-unknown system
consts, prefs, tables, samples = cn.resolve(f)
dev = torch.device("cuda", 0)
d_f = torch.from_numpy(f).to(dev); d_smp = torch.from_numpy(samples).to(dev)
d_S = torch.empty((1, cn.plan.P, cn.plan.P, pts), dtype=torch.complex128, device=dev)
d_st = torch.zeros(1, dtype=torch.int32, device=dev)
ms = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); cn.handle.run_device(d_f, consts, prefs, None, d_smp, d_S, d_st); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
ki = cn.handle.kernel_info()
print("reach %d n=%d band=%d P=%d fused=%s: class %d B=%d thr=%d regs=%d ctas=%d smem=%d -> %.3f ms, %.3e solves/s" % (
    reach, cn.plan.n, cn.plan.band, cn.plan.P, os.environ.get("HVB_BAND_FUSED"), ki["kernel_class"], ki["group"], ki["threads"], ki["regs"],
    ki["ctas_per_sm"], ki["smem_bytes"], min(ms), pts / (min(ms) * 1e-3)))
