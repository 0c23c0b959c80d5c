"""GPU debug sweep: every golden circuit through every solver mode / kernel class; prints max errors.
Run on a GPU box:  python tools/gpu_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
from heavi_b200 import plan as pl
from heavi_b200.engine import CompiledNetwork

def ulps(a, b):
    d = np.abs(a - b); s = np.maximum(np.abs(a), np.abs(b)); s[s == 0] = 1
    return float((d / s / 2.2e-16).max()) if d.size else 0.0

def check(name, m, drivers=None, values=None, Sgold=None):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    f = g["f"]
    tab = m.stamp_table()
    out = [name]
    modes = ["full"] + (["norton"] if pl.norton_legal(tab) else [])
    first = True
    for force_block in ("0", "1"):
        os.environ["HVB_FORCE_BLOCK"] = force_block
        for mode in modes:
            try:
                cn = CompiledNetwork(m, mode=mode)
                if first:
                    first = False
                    Y = cn.assemble(g["y_f"], drivers, None if values is None else values[-1:])
                    Yg = np.zeros_like(Y); Yg[g["y_rows"], g["y_cols"], g["y_fidx"]] = g["y_vals"]
                    out.append("Yulp %.1f" % ulps(Y, Yg))
                S = cn.run(f, drivers, values)
                ref = g["S"][None] if Sgold is None else Sgold
                ext = g["S_ext"][None] if Sgold is None else Sgold
                ki = cn.handle.kernel_info()
                out.append("%s/k%d.%dx%d: dref %.1e dext %.1e" % (mode, ki["kernel_class"], ki["group"], ki["pmax"],
                           np.abs(S - ref).max(), np.abs(S - ext).max()))
            except Exception as e:
                out.append("%s/blk%s: EXC %r" % (mode, force_block, e))
                if "illegal" in repr(e) or "launch failure" in repr(e):
                    print(" | ".join(out), flush=True)
                    sys.exit(3)
    print(" | ".join(out), flush=True)

if __name__ == "__main__":
    t0 = time.time()
    check("ka_divider", circuits.build_divider(hv)[0])
    check("ka_line", circuits.build_matched_line(hv)[0])
    check("c1_readme", circuits.build_c1(hv)[0])
    for b in ["bandpass", "lowpass", "highpass", "bandstop"]:
        check("c2_" + b, circuits.build_c2(hv, b)[0])
    check("c3_touchstone", circuits.build_c3(hv, nF=100)[0])
    m, f, mc = circuits.build_c4(hv)
    g = np.load(os.path.join(ROOT, "tests", "golden", "c4_mc.npz"))
    check("c4_mc", m, mc._random_numbers, g["mc_draws"], g["mc_S"])
    check("c5_ladder", circuits.build_c5(hv, nF=10)[0])
    for s in range(48):
        check("rnd_%02d" % s, circuits.build_random(hv, s)[0])
    print("done in %.1f s" % (time.time() - t0))
