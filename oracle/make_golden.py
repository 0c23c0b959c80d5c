"""Generate `tests/golden/*.npz` by running the REAL reference in this container.

    NUMBA_CACHE_DIR=/tmp/numba_cache_heavi_ref python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The reference ships no tests or golden vectors, so
parity is pinned by these files: each holds, for one circuit built by `tests/circuits.py`,

  f            float64 frequencies fed to run_sp
  kinds/off/ids   component class order and the MNA indices of every component, in the
                  reference's per-class `ids` order               (exact-match target)
  indices      the (P,4) int32 port table of network.py:400-404  (exact-match target)
  N, M
  y_f, y_rows, y_cols, y_vals   non-zeros of the reference's dense MNA tensor at a few
                  frequencies                                     (<= few ulp target)
  S            the reference's S[P,P,nF] (run_sp, zgelsd path)    (1e-12 + 1e-10|S| target)
  S_ext        the same matrices solved in 80-bit arithmetic       (adjudicator, App. E)
  f32          the float32 frequency axis run_sp returns

It also prints how far the numpy oracle (oracle/mna_oracle.py) is from the reference on each.

`mc_twoport.npz` is different in kind: the reference's own Monte-Carlo loop (`for i in mc.iterate(6)`,
np.random.seed(11)) over a netlist with lib/mc.py's RandomTwoPort -- per-trial draws and per-trial S.
`python oracle/make_golden.py --only-mc-twoport` regenerates just that file.
`optim_costs.npz` / `param_sweep.npz` (`--only-callers`): the reference's optimiser objective on a fixed
population of candidates and its ParameterSweep loop (the repeated-run_sp callers of SURVEY.md 8f).
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
from oracle.ref_shim import load_reference    # noqa: E402
import circuits                               # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def capture(name, net, f, y_points=3, extra=None):
    f = np.asarray(f, dtype=np.float64)
    recs = rx.records_from_reference(net)
    mna, Zs, idx, N, M = rx.reference_assemble(net, f)
    sp = net.run_sp(f)
    S = sp._S
    kinds, off, ids = rx.flat_ids(recs)
    pick = np.unique(np.linspace(0, len(f) - 1, y_points).astype(int))
    rows, cols, fidx, vals = [], [], [], []
    for k, fi in enumerate(pick):
        r, c = np.nonzero(mna[:, :, fi])
        rows.append(r); cols.append(c); fidx.append(np.full(r.shape, k)); vals.append(mna[r, c, fi])
    S_ext = orc.solve_extended(mna, Zs, idx, N, M).astype(np.complex128)

    # oracle vs reference, reported at generation time
    Y_or = orc.assemble(recs, N + M, f)
    dy = np.max(np.abs(Y_or - mna)) if mna.size else 0.0
    S_or = orc.solve_sp(mna, Zs, idx, N, M, method="lstsq")
    S_lu = orc.solve_sp_lu_batched(mna, Zs, idx, N, M)
    tol = 1e-12 + 1e-10 * np.abs(S)
    print("%-14s N=%3d M=%d P=%d nF=%5d | oracle: max|dY|=%.1e max|dS(lstsq)|=%.1e max|dS(lu)|=%.1e "
          "fail(lu)=%d/%d | ref-ext=%.1e lu-ext=%.1e"
          % (name, N, M, idx.shape[0], len(f), dy, np.max(np.abs(S_or - S)), np.max(np.abs(S_lu - S)),
             int(np.sum(np.any(np.abs(S_lu - S) > tol, axis=(0, 1)))), len(f),
             np.max(np.abs(S - S_ext)), np.max(np.abs(S_lu - S_ext))), flush=True)

    out = dict(f=f, kinds=kinds, off=off, ids=ids, indices=idx, N=N, M=M,
               y_f=f[pick], y_fidx=np.concatenate(fidx).astype(np.int32),
               y_rows=np.concatenate(rows).astype(np.int32), y_cols=np.concatenate(cols).astype(np.int32),
               y_vals=np.concatenate(vals), S=S, S_ext=S_ext, f32=sp.f)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def mc_twoport(hv):
    """lib/mc.py RandomTwoPort under the reference's own Monte-Carlo loop, np.random.seed(11), 6 trials."""
    m, f, mc = circuits.build_random_twoport(hv)
    np.random.seed(11)
    draws, Ss = [], []
    for _ in mc.iterate(6):
        draws.append([r._value for r in mc._random_numbers])
        Ss.append(m.run_sp(f)._S)
    np.savez_compressed(os.path.join(GOLD, "mc_twoport.npz"), f=f, mc_draws=np.array(draws), mc_S=np.stack(Ss))
    print("mc_twoport     %d trials, %d randoms, nF=%d" % (len(Ss), len(mc._random_numbers), len(f)), flush=True)


def callers(hv):
    """The reference's own repeated-`run_sp` callers: the optimiser objective (lib/optim.py:386-438, one run_sp
    per candidate) on a fixed population, and the ParameterSweep loop (numeric.py:418-446)."""
    m, opt = circuits.build_optim(hv)
    X = circuits.optim_population()
    obj = opt.generate_objective()
    costs = np.array([obj(X[:, k]) for k in range(X.shape[1])])
    m2, opt2 = circuits.build_optim(hv)
    obj_dw = opt2.generate_objective(differential_weighting=0.3)
    costs_dw = np.array([obj_dw(X[:, k]) for k in range(X.shape[1])])
    np.savez_compressed(os.path.join(GOLD, "optim_costs.npz"), X=X, costs=costs, costs_dw=costs_dw)
    print("optim_costs    %d candidates, cost range %.3g .. %.3g" % (X.shape[1], costs.min(), costs.max()), flush=True)
    m3, sweep, f = circuits.build_sweep(hv)
    for ix, vals in sweep.iterate():
        sweep.submit(m3.run_sp(f), f)
    nd = sweep.S
    np.savez_compressed(os.path.join(GOLD, "param_sweep.npz"), f=f, S=np.asarray(nd._sdata),
                        axes0=np.asarray(nd._axes[0]), axes1=np.asarray(nd._axes[1]))
    print("param_sweep    grid %s" % (np.asarray(nd._sdata).shape,), flush=True)


def main():
    hv = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "--only-mc-twoport":
        os.makedirs(GOLD, exist_ok=True)
        return mc_twoport(hv)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-callers":
        os.makedirs(GOLD, exist_ok=True)
        return callers(hv)
    os.makedirs(GOLD, exist_ok=True)
    if not os.path.exists(circuits.SYNTH8):
        circuits.write_synth8()

    m, f = circuits.build_c1(hv)
    capture("c1_readme", m, f)

    for band in ("bandpass", "lowpass", "highpass", "bandstop"):
        m, f = circuits.build_c2(hv, band)
        capture("c2_" + band, m, f[::50])          # 2001 of the 100 001 points (points are independent)

    m, f = circuits.build_c3(hv, nF=1_000_000)
    capture("c3_touchstone", m, f[::3907][:256], y_points=2)

    # config 4: first 4 Monte-Carlo samples, np.random.seed(7), 401 of the 2001 points
    m, f, mc = circuits.build_c4(hv)
    np.random.seed(7)
    fsub = f[::5]
    draws, Ss = [], []
    for _ in mc.iterate(4):
        draws.append([r._value for r in mc._random_numbers])
        Ss.append(m.run_sp(fsub)._S)
    capture("c4_mc", m, fsub, extra=dict(mc_draws=np.array(draws), mc_S=np.stack(Ss)))

    m, f = circuits.build_c5(hv, nF=1_000_000)
    capture("c5_ladder", m, f[::66667][:12], y_points=1)

    mc_twoport(hv)
    callers(hv)

    m, f = circuits.build_divider(hv)
    capture("ka_divider", m, f)
    m, f = circuits.build_matched_line(hv)
    capture("ka_line", m, f)

    for seed in range(48):
        m, f = circuits.build_random(hv, seed)
        capture("rnd_%02d" % seed, m, f, y_points=len(f))


if __name__ == "__main__":
    main()
