"""Randomised parity sweep of the banded kernels against the CPU oracle (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heavi_b200 as hv
import circuits
from heavi_b200.engine import CompiledNetwork
from oracle import mna_oracle as orc
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst, n_band = 0.0, 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 60):
    reach = int(rng.integers(1, 9)); K = int(rng.integers(max(9, 7 * reach), 220)); nports = int(rng.integers(1, 9))
    mode = [None, None, "full"][int(rng.integers(0, 3))]
    m, f = circuits.build_band_ladder(hv, K=K, reach=reach, nports=nports, seed=int(rng.integers(0, 10**6)), nF=7)
    cn = CompiledNetwork(m, mode=mode)
    plan, handle = cn.variant(10**9)                      # the variant a large batch would use
    ki = handle.kernel_info()
    S = handle.run_host(f, *cn.resolve(f, plan=plan))[0][0]
    So = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f, method="lu_batched")
    err = float(np.max(np.abs(S - So) / (1e-12 + 1e-10 * np.abs(So))))
    worst = max(worst, err); n_band += ki["kernel_class"] == 2
    flag = "" if err <= 1.0 else "   <-- OUT OF TOLERANCE"
    print("reach %d K %3d P %d mode %-5s n %3d band %2d class %d B %d: err/tol %.2e%s" % (reach, K, nports, mode, plan.n, plan.band, ki["kernel_class"], ki["group"], err, flag), flush=True)
print("worst err/tol %.3e, banded kernel used in %d cases" % (worst, n_band))
