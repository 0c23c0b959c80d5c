"""GPU (-m gpu): kernel class 3 -- the kernel the library generates for a chain-like netlist and compiles
with NVRTC (csrc/hvb_gen.cpp, hvb_jit.cpp) -- through the public API / C-ABI against the oracle, the
reference goldens, the precompiled kernels, and size-independent properties at BASELINE sizes."""
import os
import warnings

import numpy as np
import pytest

import circuits
import heavi_b200 as hv
from heavi_b200.engine import CompiledNetwork
from helpers import golden, parity_report
from oracle import mna_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gen_on():
    keep = {k: os.environ.get(k) for k in ("HVB_NO_GEN", "HVB_GEN_MODE")}
    os.environ.pop("HVB_NO_GEN", None)
    yield
    for k, v in keep.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def oracle_S(m, f):
    return orc.run_sp(m.records(), m.n_nodes, len(m.sources), np.asarray(f, dtype=np.float64), method="lu_batched")


def library_variant(m):
    os.environ["HVB_NO_GEN"] = "1"
    try:
        return CompiledNetwork(m)
    finally:
        os.environ.pop("HVB_NO_GEN")


@pytest.mark.parametrize("gen_mode", ["1", "2", "3"])
@pytest.mark.parametrize("name,build", [("c1_readme", lambda: circuits.build_c1(hv)[0]),
                                        ("c2_bandpass", lambda: circuits.build_c2(hv, "bandpass")[0]),
                                        ("c2_lowpass", lambda: circuits.build_c2(hv, "lowpass")[0]),
                                        ("c2_highpass", lambda: circuits.build_c2(hv, "highpass")[0]),
                                        ("c5_ladder", lambda: circuits.build_c5(hv, nF=16)[0])])
def test_configs_on_the_generated_kernel_match_reference_golden(name, build, gen_mode):
    os.environ["HVB_GEN_MODE"] = gen_mode               # 1 straight-line code, 2 loop over row shapes, 3 the loop walked from both ends
    g = golden(name)
    m = build()
    S = m.run_sp(g["f"])
    ki = m._compiled().handle.kernel_info()
    assert ki["kernel_class"] == 3 and ki["gen_rolled"] == (gen_mode != "1") and ki["gen_two_sided"] == (gen_mode == "3"), ki
    assert S.f.dtype == np.float32 and np.array_equal(S.f, g["f32"])
    frac, adjudicated, ext = parity_report(S._S, g["S"], g["S_ext"])
    assert frac >= 0.998 and adjudicated and ext < 5e-14, (frac, adjudicated, ext)


@pytest.mark.parametrize("n_sections,n_ports", [(3, 1), (3, 3), (7, 2), (16, 5), (33, 4), (64, 16), (200, 8), (501, 3), (2001, 2)])
def test_chains_of_every_size_against_oracle_and_library_kernels(n_sections, n_ports):
    rng = np.random.default_rng(n_sections)
    m = hv.Model(suppress_loadbar=True)
    nodes = [m.node() for _ in range(n_sections)]
    for t in sorted(set(int(round(i * (n_sections - 1) / max(n_ports - 1, 1))) for i in range(n_ports))):
        m.new_port(50.0, nodes[t])
    for k in range(n_sections - 1):
        if k % 4 == 1:
            m.transmissionline(nodes[k], nodes[k + 1], float(rng.uniform(30, 90)), float(rng.uniform(2, 4)), float(rng.uniform(1e-3, 9e-3)))
        elif k % 4 == 3:
            m.resistor(nodes[k], nodes[k + 1], float(rng.uniform(5, 80)))
        else:
            m.inductor(nodes[k], nodes[k + 1], float(10 ** rng.uniform(-9.3, -8.3)))
    for k in range(n_sections):
        m.capacitor(nodes[k], m.gnd, float(10 ** rng.uniform(-12.5, -11.5)))
    f = np.linspace(0.3e9, 6e9, 257)
    S = m.run_sp(f)._S
    assert m._compiled().handle.kernel_info()["kernel_class"] == 3
    sub = slice(0, None, 16)
    assert np.max(np.abs(S[..., sub] - oracle_S(m, f[sub]))) < 1e-12
    lib = library_variant(m)
    assert lib.handle.kernel_info()["kernel_class"] != 3
    assert np.max(np.abs(S - lib.run(f)[0])) < 1e-12


def test_two_sided_chain_kernel_at_size_and_its_refusal_path():
    """configs[4]'s netlist: long chains are walked from both ends (csrc/hvb_gen.cpp); against the oracle, against the
    one-sided kernel of the same plan, and -- on a weakly coupled resonator chain whose cut is unsafe at the halves'
    resonances -- the library's own repeat on the one-sided kernel, on every host entry."""
    m, f = circuits.build_c5(hv, nF=300_000)
    S = m.run_sp(f)._S
    cn = m._compiled()
    ki = cn.handle.kernel_info()
    assert ki["kernel_class"] == 3 and ki["gen_rolled"] == 1 and ki["gen_two_sided"] == 1, ki
    idx = np.arange(0, 300_000, 7499)
    assert np.max(np.abs(S[..., idx] - oracle_S(m, f[idx]))) < 1e-12
    assert np.max(np.abs(S - S.transpose(1, 0, 2))) < 1e-12                      # reciprocal network, equal port impedances
    cn.handle.set_tunable("gen_two_sided", 0)
    S1 = m.run_sp(f)._S
    assert cn.handle.kernel_info()["gen_two_sided"] == 0
    assert np.max(np.abs(S - S1)) < 1e-13
    cn.handle.set_tunable("gen_two_sided", 1)
    assert np.array_equal(m.run_sp(f)._S, S) and cn.handle.kernel_info()["gen_two_sided"] == 1

    # device-resident entry, whole job in one launch: a launch of a few waves gets the 128-register budget (512 threads per SM;
    # csrc/hvb_api.cu, ensure_gen_kernel), a deep one the compiler's own 168 registers (384 threads); same arithmetic either way
    Sa = m.run_sp_resident(f[:125_000]).cpu().numpy()[0]
    ka = cn.handle.kernel_info()
    Sb = m.run_sp_resident(f).cpu().numpy()[0]
    kb = cn.handle.kernel_info()
    m2, f2 = circuits.build_c5(hv, nF=600_000)
    Sc = m2.run_sp_resident(f2[:600_000]).cpu().numpy()[0]
    kc = m2._compiled().handle.kernel_info()
    assert ka["gen_two_sided"] == kb["gen_two_sided"] == kc["gen_two_sided"] == 1
    assert ka["threads"] * ka["ctas_per_sm"] == kb["threads"] * kb["ctas_per_sm"] == 512 and kc["threads"] * kc["ctas_per_sm"] == 384, (ka, kb, kc)
    assert ka["regs"] <= 128 < kc["regs"] <= 168
    assert np.max(np.abs(Sa - S[..., :125_000])) < 1e-13 and np.max(np.abs(Sb - S)) < 1e-13
    idx2 = np.arange(0, 600_000, 14999)
    assert np.max(np.abs(Sc[..., idx2] - oracle_S(m2, f2[idx2]))) < 1e-12

    fr = np.linspace(0.2e9, 8e9, 4000)
    ref = None
    os.environ["HVB_GEN_MODE"] = "3"          # ports at the two ends only: auto mode would not cut this chain (no saving)
    for entry in ("run_sp", "run_sp_async", "run_sp_resident"):
        mr = circuits.build_lc_chain(hv, 44, None, coupling=1e-15)
        if ref is None:
            ref = oracle_S(mr, fr)
        with warnings.catch_warnings():
            warnings.simplefilter("error")                                       # no singular-system warning: the system is fine
            if entry == "run_sp":
                Sr = mr.run_sp(fr)._S
            elif entry == "run_sp_async":
                Sr = mr.run_sp_async(fr).result()._S
            else:
                Sr = mr.run_sp_resident(fr).cpu().numpy()[0]
        ki = mr._compiled().handle.kernel_info()
        assert ki["kernel_class"] == 3 and ki["gen_rolled"] == 1 and ki["gen_two_sided"] == 0, (entry, ki)   # went back for good
        assert np.all(np.abs(Sr - ref) <= 1e-12 + 1e-10 * np.abs(ref)), entry


def test_full_size_monte_carlo_and_sweep_properties():
    """configs[3] per-GPU share (1250 samples x 2001 points) and a 1 M-point configs[1] sweep: reciprocity,
    losslessness of the LC ladder (S^H S = I), strided oracle check, tiling-independence."""
    m, f, mc = circuits.build_c4(hv)
    np.random.seed(7)
    values = mc.draw(1250)
    S = m.run_sp_batch(f, mc._random_numbers, values)
    assert S.shape == (1250, 2, 2, 2001) and m._compiled().handle.kernel_info()["kernel_class"] == 3
    assert np.max(np.abs(S[:, 0, 1] - S[:, 1, 0])) < 1e-12
    SS = np.einsum("sjif,sjkf->sikf", S.conj(), S)
    assert np.max(np.abs(SS - np.eye(2)[None, :, :, None])) < 5e-11
    for s in (0, 611, 1249):
        for r, v in zip(mc._random_numbers, values[s]):
            r._value = float(v)
        assert np.max(np.abs(S[s][..., ::97] - oracle_S(m, f[::97]))) < 1e-12
    cn = m._compiled()
    cn.handle.set_tunable("chunk_mb", 1)
    assert np.array_equal(m.run_sp_batch(f, mc._random_numbers, values[:100]), S[:100])
    m2, f2 = circuits.build_c2(hv, "bandpass", nF=1_000_001)
    S2 = m2.run_sp(f2)._S
    assert np.max(np.abs(S2[0, 1] - S2[1, 0])) < 1e-12
    idx = np.arange(0, 1_000_001, 9973)
    assert np.max(np.abs(S2[..., idx] - oracle_S(m2, f2[idx]))) < 1e-12
    assert np.array_equal(S2[..., idx], m2.run_sp(f2[idx])._S)       # a point does not depend on its neighbours


def test_status_bits_statistics_and_resident_entry():
    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    mid = m.node()
    m.inductor(a, mid, 1e-9)
    m.inductor(mid, b, 1e-9)
    assert m._compiled().handle.kernel_info()["kernel_class"] == 3
    with pytest.raises(np.linalg.LinAlgError):                      # like numba's lstsq on NaN/Inf
        m.run_sp(np.array([0.0, 1e9]))
    m2 = hv.Model(suppress_loadbar=True)
    p, q = m2.quick_port(50), m2.quick_port(50)
    x = m2.node()
    m2.resistor(p, x, 10.0)
    m2.admittance(x, q, 0.0)                                        # port 2 floats behind a zero admittance
    m2.admittance(q, m2.gnd, 0.0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        S = m2.run_sp(np.array([1e9]))
    assert np.isfinite(S._S).all()                                  # the port conductance keeps it regular
    m3, f3, mc = circuits.build_c4(hv, nF=301)
    np.random.seed(11)
    full = m3.run_mc(f3, mc, 23)
    np.random.seed(11)
    dev = m3.run_mc_stats(f3, mc, 23)
    assert np.max(np.abs(full.S(2, 1).amplitude_mean - dev.S(2, 1).amplitude_mean)) < 1e-12
    import torch
    Sr = m3._compiled().run_resident(f3)
    assert Sr.is_cuda and np.array_equal(Sr.cpu().numpy()[0], m3.run_sp(f3)._S)


def test_more_than_64_constant_sets_on_one_plan_evict_instead_of_failing():
    """The generated code carries the run's constants, one compiled variant per distinct set; a plan keeps 64 of them and
    drops the oldest beyond that (csrc/hvb_api.cu, ensure_gen_kernel).  Every set must give the result a fresh model with
    those values gives, and a set that was evicted comes back bit-identical."""
    m, f = circuits.build_c1(hv, nF=257)
    cn = m._compiled()
    consts, prefs, tables, samples = cn.resolve(f, None, None)
    assert cn.handle.kernel_info()["kernel_class"] == 3 and len(consts) > 4 and consts[4] == 8e-10   # the SMD part's series inductance
    first = None
    for k in range(70):
        c = consts.copy()
        c[4] = consts[4] * (1.0 + 1e-2 * k)                      # one component value, 70 distinct constant sets
        S, status = cn.handle.run_host(f, c, prefs, tables, samples)
        assert status == 0
        if k == 0:
            first = S.copy()
        if k in (1, 33, 69):
            S2, _ = cn.handle.run_host(f, c, prefs, tables, samples)
            assert np.array_equal(S, S2)
            assert np.max(np.abs(S - first)) > 1e-9                  # the changed value did reach the kernel
    S0, status = cn.handle.run_host(f, consts, prefs, tables, samples)  # set 0 was evicted (64 newer ones since)
    assert status == 0 and np.array_equal(S0, first)
    assert np.max(np.abs(S0[0] - oracle_S(m, f))) < 1e-12


def test_parameter_layout_change_regenerates_the_kernel():
    """A callable that returns an array is a table, one that returns a scalar becomes a constant for that
    run (resolve_run): the value formulas are specialised on that layout, so the library keeps one kernel
    per layout."""
    state = {"scalar": False}

    def z(fr):
        return 5.0 + 0.0j if state["scalar"] else 5.0 + 1e-9j * fr

    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    mid = m.node()
    m.impedance(a, mid, z)
    m.inductor(mid, b, 2e-9)
    m.capacitor(mid, m.gnd, 1e-12)
    f = np.linspace(1e9, 3e9, 41)
    S1 = m.run_sp(f)._S.copy()
    assert np.max(np.abs(S1 - oracle_S(m, f))) < 1e-13
    state["scalar"] = True
    S2 = m.run_sp(f)._S.copy()
    assert np.max(np.abs(S2 - oracle_S(m, f))) < 1e-13 and not np.array_equal(S1, S2)
    state["scalar"] = False
    assert np.array_equal(m.run_sp(f)._S, S1)


# ---------------------------------------------------------------------------------------------------
# kernel E: hub netlists (N-port S block + matching chains), csrc/hvb_gen_hub.cpp
# ---------------------------------------------------------------------------------------------------
def test_config3_runs_on_the_hub_kernel_and_matches_reference_golden():
    g = golden("c3_touchstone")
    m, f = circuits.build_c3(hv, nF=200_000)
    S = m.run_sp(g["f"])
    ki = m._compiled().handle.kernel_info()
    assert ki["kernel_class"] == 3 and ki["gen_hub"] == 1 and ki["threads"] == 32, ki
    frac, adjudicated, ext = parity_report(S._S, g["S"], g["S_ext"])
    assert frac == 1.0 and adjudicated and ext < 5e-14
    # BASELINE size property checks: reciprocity of the synthetic block, strided oracle check, precompiled kernel
    Sb = m.run_sp(f)._S
    assert np.max(np.abs(Sb - np.swapaxes(Sb, 0, 1))) < 1e-11
    idx = np.arange(0, 200_000, 4001)
    assert np.max(np.abs(Sb[..., idx] - oracle_S(m, f[idx]))) < 1e-12
    lib = library_variant(m)
    assert lib.handle.kernel_info()["kernel_class"] == 0
    assert np.max(np.abs(Sb[..., idx] - lib.run(f[idx])[0])) < 1e-12


@pytest.mark.parametrize("n_block,chain_rows,port_on", [
    (4, (2, 1, 0, 3), (0, 0, None, 1)), (2, (1, 1), (0, 0)), (3, (0, 2, 0), (0, 1, 0)), (8, (1,) * 8, (0,) * 8),
    (5, (4, 0, 2, 0, 1), (2, None, 0, 0, None))])
def test_hub_structures_against_oracle_and_library_kernels(n_block, chain_rows, port_on):
    m, f = circuits.build_hub(hv, n_block, chain_rows, port_on, seed=n_block, nF=129)
    S = m.run_sp(f)._S
    ki = m._compiled().handle.kernel_info()
    assert ki["kernel_class"] == 3 and ki["gen_hub"] == 1
    assert np.max(np.abs(S - oracle_S(m, f))) < 1e-12
    assert np.max(np.abs(S - library_variant(m).run(f)[0])) < 1e-12


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_hub_kernel_inversions_exchange_rows_when_they_must(seed):
    """S -> 0.3 S - 0.995 I (tests/test_codegen.py, same case on the CPU): both in-place inversions of kernel E pivot at
    most points; the result must not care."""
    m, f = circuits.build_hub(hv, 6, (1, 2, 0, 1, 0, 2), (0, 1, 0, 0, None, 0), seed=seed, nF=513, scale=0.3, diag_bias=0.995)
    S = m.run_sp(f)._S
    ki = m._compiled().handle.kernel_info()
    assert ki["kernel_class"] == 3 and ki["gen_hub"] == 1 and ki["smem_bytes"] == 6 * 6 * 32 * 16
    ref = oracle_S(m, f)
    assert np.all(np.abs(S - ref) <= 1e-12 + 1e-10 * np.abs(ref)), np.max(np.abs(S - ref))


def test_hub_kernel_unsafe_pivot_falls_back_to_the_dense_kernel():
    m = hv.Model(suppress_loadbar=True)
    b, c, mid = m.node(), m.node(), m.node()
    m.new_port(50.0, b)
    m.new_port(50.0, c)
    C2, f0 = 2e-12, 1e9
    m.capacitor(mid, b, C2)
    m.inductor(mid, m.gnd, 1.0 / ((2 * np.pi * f0) ** 2 * C2))
    m.n_port_S(m.gnd, [b, c], [[lambda f: 0.1 + 0 * f, lambda f: 0.7 + 0 * f], [lambda f: 0.7 + 0 * f, lambda f: 0.2 + 0 * f]], 50.0)
    f = np.array([0.5e9, f0, 1.7e9])
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        S = m.run_sp(f)._S
    cn = m._compiled()
    assert cn._hub and cn._dense_fallback is not None              # the flagged call was repeated on the dense kernel
    assert np.max(np.abs(S - oracle_S(m, f))) < 1e-12


def test_missing_nvrtc_falls_back_to_the_precompiled_kernels():
    """No NVRTC (a stripped-down deployment): the netlist still runs on the GPU, on the library's own kernels."""
    import subprocess
    import sys
    code = (
        "import os, sys, warnings; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "os.environ['HVB_NVRTC_LIB'] = '/nonexistent/libnvrtc.so'\n"
        "import ctypes, ctypes.util\n"
        "import numpy as np, heavi_b200 as hv, circuits\n"
        "from oracle import mna_oracle as orc\n"
        "m, f = circuits.build_c2(hv, 'bandpass', nF=501)\n"
        "with warnings.catch_warnings(record=True) as w:\n"
        "    warnings.simplefilter('always')\n"
        "    S = m.run_sp(f)._S\n"
        "ref = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f, method='lu_batched')\n"
        "print('FALLBACK' if any('precompiled' in str(x.message) for x in w) else 'GENERATED', m._compiled().handle.kernel_info()['kernel_class'], float(np.max(np.abs(S - ref))))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("HVB_NO_GEN", None)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    kind, kclass, err = out.stdout.strip().split()[-3:]
    assert kind == "FALLBACK" and float(err) < 1e-12 and int(kclass) in (0, 2)
