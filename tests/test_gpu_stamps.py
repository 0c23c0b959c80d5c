"""GPU (-m gpu): a plan created by the library's own stamp-table compiler (`hvb_plan_create_from_stamps`, the path a
host without Python takes: consts / prefs stay NULL) against the reference goldens and the Python-compiled path."""
import numpy as np
import pytest

import circuits
import heavi_b200 as hv
from heavi_b200.engine import StampPlanHandle
from helpers import golden, parity_report, registry

pytestmark = pytest.mark.gpu
REG = registry(hv)
NAMES = ["c1_readme", "c2_bandpass", "c2_lowpass", "c2_highpass", "c2_bandstop", "c3_touchstone", "c5_ladder", "ka_divider", "ka_line"] + \
        ["rnd_%02d" % k for k in range(0, 48, 3)]


@pytest.mark.parametrize("name", NAMES)
def test_plans_compiled_below_the_abi_match_reference_golden(name, kernel_path):
    g = golden(name)
    m = REG[name]()
    h = StampPlanHandle(m.stamp_table())
    S, status = h.run(g["f"])
    assert status == 0
    frac, adjudicated, ext = parity_report(S[0], g["S"], g["S_ext"])
    assert frac >= 0.998 and adjudicated and ext < 5e-14, (frac, adjudicated, ext)
    ref = m.run_sp(g["f"])._S                                   # the Python-compiled plan picks the same kernel family
    assert h.kernel_info()["kernel_class"] == m._compiled().handle.kernel_info()["kernel_class"]
    assert np.max(np.abs(S[0] - ref)) < 1e-12


def test_monte_carlo_samples_through_the_stamp_path(kernel_path):
    m, f, mc = circuits.build_c4(hv, nF=301)
    np.random.seed(5)
    values = mc.draw(17)
    h = StampPlanHandle(m.stamp_table())
    S, status = h.run(f, mc._random_numbers, values)
    assert status == 0 and np.array_equal(S, m.run_sp_batch(f, mc._random_numbers, values))
