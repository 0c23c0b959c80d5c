"""Run the REAL reference and the numpy oracle side by side, in this container, on circuits that are
NOT in `tests/golden/` (fresh seeds, other sizes), and report / assert how far apart they are.

    NUMBA_CACHE_DIR=/tmp/numba_cache_heavi_ref python oracle/check_oracle_vs_ref.py [--seeds 100:140]

TEST INFRASTRUCTURE ONLY (the product never imports anything from `oracle/`).  The golden files pin the
oracle on a fixed set of circuits on every machine; this script is the open-ended version of the same check
for a machine that has `/root/reference`: stamp index tables must match exactly, the assembled MNA tensor to a
few ulp, and S within `1e-12 + 1e-10 |S|` of the reference's `run_sp` (network.py:386-426, zgelsd path),
outliers adjudicated against the 80-bit solve of the reference's own matrices (`mna_oracle.solve_extended`).
Exit status 1 on any violation.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import mna_oracle as orc          # noqa: E402
from oracle import ref_extract as rx          # noqa: E402
from oracle.ref_shim import load_reference, reference_available    # noqa: E402
import circuits                               # noqa: E402


def compare(name, net, f):
    """One circuit: reference records -> oracle assembly and solve, against the reference's own tensors."""
    f = np.asarray(f, dtype=np.float64)
    recs = rx.records_from_reference(net)
    mna, Zs, idx, N, M = rx.reference_assemble(net, f)
    S_ref = net.run_sp(f)._S
    Y_or = orc.assemble(recs, N + M, f)
    scale = max(1.0, float(np.max(np.abs(mna)))) if mna.size else 1.0
    dy = float(np.max(np.abs(Y_or - mna))) / scale if mna.size else 0.0
    S_or = orc.run_sp(recs, N, M, f, method="lu_batched")
    S_ext = orc.solve_extended(mna, Zs, idx, N, M).astype(np.complex128)
    tol = 1e-12 + 1e-10 * np.abs(S_ref)
    bad = np.abs(S_or - S_ref) > tol
    # a point outside the gate is the reference's SVD noise if the oracle is the one that agrees with 80-bit
    unexplained = bad & (np.abs(S_or - S_ext) > 5e-14 + 1e-12 * np.abs(S_ext))
    ok = dy < 1e-14 and not unexplained.any()
    print("%-22s N=%3d M=%d P=%d nF=%5d  rel|dY|=%.1e  max|S_or-S_ref|=%.1e  outside gate=%d  unexplained=%d  "
          "|S_or-S_80bit|=%.1e |S_ref-S_80bit|=%.1e  %s"
          % (name, N, M, idx.shape[0], len(f), dy, float(np.max(np.abs(S_or - S_ref))), int(bad.sum()),
             int(unexplained.sum()), float(np.max(np.abs(S_or - S_ext))), float(np.max(np.abs(S_ref - S_ext))),
             "ok" if ok else "FAIL"), flush=True)
    return ok


def main(argv):
    if not reference_available():
        print("reference tree not present: nothing to check (the golden files pin the oracle instead)")
        return 0
    lo, hi = 100, 124
    if "--seeds" in argv:
        lo, hi = (int(x) for x in argv[argv.index("--seeds") + 1].split(":"))
    hv = load_reference()
    if not os.path.exists(circuits.SYNTH8):
        circuits.write_synth8()
    ok = True
    for seed in range(lo, hi):                              # the golden set is seeds 0..47
        m, f = circuits.build_random(hv, seed)
        ok &= compare("rnd_%03d" % seed, m, f)
    m, f = circuits.build_c1(hv, nF=333)
    ok &= compare("c1 nF=333", m, f)
    for band in ("bandpass", "lowpass", "highpass", "bandstop"):
        m, f = circuits.build_c2(hv, band, nF=777)
        ok &= compare("c2 %s nF=777" % band, m, f)
    m, f = circuits.build_c3(hv, nF=1_000_000)
    ok &= compare("c3 other stride", m, f[1234::9973][:64])
    m, f = circuits.build_c5(hv, nF=1_000_000)
    ok &= compare("c5 other stride", m, f[777::99991][:8])
    for K in (7, 33):
        m, f = circuits.build_c5(hv, nF=16, K=K)
        ok &= compare("ladder K=%d" % K, m, f)
    for seed in (3, 4):
        m, f = circuits.build_band_ladder(hv, K=24, reach=2, nports=3, seed=seed, nF=12)
        ok &= compare("band ladder seed %d" % seed, m, f)
        m, f = circuits.build_twoport_cascade(hv, sections=5, nF=17, seed=seed)
        ok &= compare("two-port cascade %d" % seed, m, f)
        m, f = circuits.build_hub(hv, seed=seed)
        ok &= compare("hub seed %d" % seed, m, f)
    print("oracle vs reference:", "all within the gate" if ok else "VIOLATIONS")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
