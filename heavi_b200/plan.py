"""Freeze a netlist into the stamp table / device program the CUDA kernels interpret.

Two layers:

`StampTable` -- the *structural* view of the circuit, one entry per component in creation order:
kind, MNA indices in the reference's per-class `ids` order (ground = 0, duplicates kept), and the
parameters.  Plus `port_table[P,4]` = exactly the `indices` array of network.py:400-404, `N`, `M`.
This is the part that must match the reference bit for bit.

`DevicePlan` -- what the C-ABI takes (`include/heavi_b200.h`, `hvb_plan_desc`).  For every point
(sample s, frequency f) the kernels run the same small program:

  1. *value program*: evaluate every frequency/sample dependent stamp value of every component
     (a capacitor's j w C, the nine entries of a line's indefinite block, the (P+1)^2 entries of an
     N-port block ...) into a per-point value array.  Values that do not depend on f or s are
     evaluated once, here, with numpy -- the reference's own arithmetic -- and shipped as constants.
  2. *cell program*: a CSR list per matrix cell (row, col) of `(value id, sign)` pairs in component
     order.  Summing a cell's list reproduces the reference's sequence of `+=` on that cell,
     including numpy's last-write-wins rule for a component that repeats an index (App. A).
     The right-hand sides (one column per port) are extra columns of the same matrix.
  3. solve [A | B] and turn node voltages into S (solvers/spfn.py:159-176) using `port_rows`.

Modes: "full" = the reference's system (ground row/column dropped, n = N+M-1, unit RHS on the source
rows); "norton" = port internal nodes and source rows eliminated analytically (n = N-1-P), legal
when nothing but the port touches its internal node; "dump" = the whole (N+M)x(N+M) tensor incl.
ground and no RHS, used by `hvb_assemble` to compare against the reference's `mna_matrix`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ---- constants shared with csrc/plan_format.h (keep in sync) -------------------------------------
PK_CONST, PK_SAMPLE, PK_SAMPLE_NEG, PK_SAMPLE_INV, PK_TABLE, PK_SPLINE, PK_BETA, PK_SMD, PK_BETA_MUL = range(9)
OP_COPY, OP_RECIP, OP_CAP, OP_IND, OP_TLINE = range(5)
COOP_NPORT_S, COOP_NPORT_Y = 0, 1
OP_WORDS = 8          # int32 words per simple op
COOP_WORDS = 8        # int32 words per cooperative op
ROW_GROUND, ROW_VIRTUAL_INT = -1, -2
MAX_NPORT = 32

_C0 = 299792458
_TWOPI = np.float64(2 * np.pi)


# ==================================================================================================
# structural stamp table
# ==================================================================================================
@dataclass
class StampTable:
    kinds: list            # component kind per component, creation order
    ids: list              # list of int lists (reference `ids` order)
    params: list           # list of dicts (SimParams / callables / numbers)
    port_table: np.ndarray # (P,4) int32: (iport-1, sig, int, gnd)
    port_src: np.ndarray   # (P,) int32 source row of each port
    port_z0: list          # SimParam per port, port_table order
    N: int
    M: int

    @property
    def P(self) -> int:
        return self.port_table.shape[0]

    @classmethod
    def from_network(cls, net) -> "StampTable":
        kinds, ids, params = [], [], []
        for comp in net.components:
            kinds.append(comp.kind)
            ids.append([int(i) for i in comp.stamp_ids()])
            params.append(comp.stamp_params())
        rows, srcs, z0 = [], [], []
        for iport, port in net.ports.items():
            rows.append((iport - 1, port.sig_node.index, port.int_node.index, port.gnd_node.index))
            srcs.append(port.sources[0].index)
            z0.append(port.Z0)
        pt = np.array(rows, dtype=np.int32).reshape(len(rows), 4)
        return cls(kinds, ids, params, pt, np.array(srcs, dtype=np.int32), z0, net.n_nodes, len(net.sources))

    def flat(self):
        """kinds / offsets / flat ids -- the integer arrays the golden files hold."""
        off = np.cumsum([0] + [len(i) for i in self.ids]).astype(np.int32)
        flat = np.array([i for lst in self.ids for i in lst], dtype=np.int32)
        return np.array(self.kinds), off, flat


# ==================================================================================================
# device plan
# ==================================================================================================
@dataclass
class DevicePlan:
    mode: str
    n: int                      # system dimension (rows)
    ncols: int                  # n + number of RHS columns (or n for "dump")
    P: int
    NM: int
    consts: np.ndarray          # complex128 [nConst]; slots >= n_static_consts are filled per run
    prefs: np.ndarray           # int32 [nPref,2] (kind, idx); table kinds may be rewritten per run
    const_vals: np.ndarray      # complex128 [nVC] plan-time stamp values (id 0 is 1+0j)
    n_dyn: int                  # number of per-point values
    ops: np.ndarray             # int32 [nOps, OP_WORDS]
    coops: np.ndarray           # int32 [nCoop, COOP_WORDS]
    cell_ptr: np.ndarray        # int32 [n*ncols+1]
    cell_ent: np.ndarray        # int32 [nnz]  (value id << 1) | negate
    row_ptr: np.ndarray         # int32 [n+1]  per-row entry lists (kernel A's assembly format)
    row_ent: np.ndarray         # int32 [nnz]  negate | value id << 1 | column << 20 | last-of-cell << 31
    port_rows: np.ndarray       # int32 [P,3] sig/int/gnd system rows, ROW_GROUND / ROW_VIRTUAL_INT
    port_zpref: np.ndarray      # int32 [P] pref index of Z0 per port
    # splines
    spl_t: np.ndarray           # float64 knots, all sets concatenated
    spl_c: np.ndarray           # complex128 coefficients, all traces concatenated
    spl_trace: np.ndarray       # int32 [nTrace,4]: t_off, nt, c_off, k
    spl_fill: np.ndarray        # complex128 [nTrace,2]
    spl_range: np.ndarray       # float64 [nTrace,2]
    # host-side bookkeeping (not shipped)
    sample_owners: list = field(default_factory=list)     # objects whose `_value` feeds a sample column
    table_params: list = field(default_factory=list)      # (pref index, const slot, callable)
    n_static_consts: int = 0
    coop_scratch: int = 0       # complex128 scratch words a cooperative op needs per point
    # fast paths (kernel A): value ops whose parameter is a real constant or a plain sample column,
    # and the closed-form epilogue S_ji = (2 V_sig_j - delta_ij) sqrt(Zi/Zj) of a Norton plan whose
    # ports are ground-referenced with constant real Z0
    fop_code: np.ndarray = None # int32 [nFast]  bits 0-1 op (COPY/RECIP/CAP/IND), bit 2: argument is a sample column
    fop_out: np.ndarray = None  # int32 [nFast]  per-point value slot
    fop_arg: np.ndarray = None  # float64 [nFast] constant, or the sample column
    epi_simple: int = 0
    epi_port_of_row: np.ndarray = None   # int32 [n]  port whose signal node is this row, else -1
    epi_scale: np.ndarray = None         # float64 [P,P]  sqrt|Zi| / sqrt|Zj| at [j,i]
    n_real: int = 0             # unknowns before identity padding (the dimension actually needed)
    max_nport: int = 0          # largest N-port block (a cooperative op needs that many lanes)
    band: int = -1              # half bandwidth after the RCM row ordering (-1: natural order kept)

    @property
    def n_tables(self) -> int:
        return len(self.table_params)

    @property
    def n_samples_cols(self) -> int:
        return len(self.sample_owners)

    def flops_per_solve(self) -> float:
        """Algorithmic real flops of one point: complex LU + P substitutions on the dimension
        actually factored."""
        n, P = self.n_real, self.P
        return (8.0 / 3.0) * n ** 3 + 8.0 * n * n * P


class _Builder:
    def __init__(self):
        self.consts: list[complex] = []
        self.prefs: list[tuple[int, int]] = []
        self.sample_owners: list = []
        self.table_params: list = []
        self.const_vals: list[complex] = [1.0 + 0.0j]      # value id 0 == 1
        self.dyn_count = 0
        self.ops: list[list[int]] = []
        self.coops: list[list[int]] = []
        self.spl_t: list[np.ndarray] = []
        self.spl_sets: dict = {}
        self.spl_c: list[np.ndarray] = []
        self.spl_trace: list[tuple[int, int, int, int]] = []
        self.spl_fill: list[tuple[complex, complex]] = []
        self.spl_range: list[tuple[float, float]] = []
        self._t_len = 0
        self._c_len = 0
        self.coop_scratch = 0

    # ---- parameters ------------------------------------------------------------------------
    def const(self, v) -> int:
        self.consts.append(complex(v))
        return len(self.consts) - 1

    def pref(self, kind: int, idx: int) -> int:
        self.prefs.append((kind, idx))
        return len(self.prefs) - 1

    def describe(self, p):
        """Normalise anything parameter-like into a description tuple (see params.py)."""
        if hasattr(p, "describe"):
            return p.describe()
        return describe_foreign(p)

    def param(self, p) -> tuple[int, tuple]:
        """-> (pref index, description).  Table parameters get a const slot they may fall back to
        when the callable returns a scalar at run time."""
        d = self.describe(p)
        tag = d[0]
        if tag == "const":
            return self.pref(PK_CONST, self.const(d[1])), d
        if tag == "sample":
            owner, op = d[1], d[2]
            for col, o in enumerate(self.sample_owners):
                if o is owner:
                    break
            else:
                self.sample_owners.append(owner)
                col = len(self.sample_owners) - 1
            return self.pref({"id": PK_SAMPLE, "neg": PK_SAMPLE_NEG, "inv": PK_SAMPLE_INV}[op], col), d
        if tag == "beta":
            # the device multiplies (2 pi f / c0) by sqrt(er): the root does not depend on f, so numpy
            # takes it here, once, exactly as `np.sqrt(er(f))` does per element (network.py:627-631)
            er = complex(d[1])
            with np.errstate(invalid="ignore"):
                root = np.sqrt(np.float64(er.real)) if er.imag == 0 else np.sqrt(np.complex128(er))
            return self.pref(PK_BETA, self.const(complex(root))), d
        if tag == "spline":
            return self.pref(PK_SPLINE, self.add_spline(d[1])), d
        if tag == "beta_mul":
            return self.pref(PK_BETA_MUL, self.const(float(d[1]))), d
        if tag == "smd":
            # closed form of a surface-mount part: four consecutive constants [kind, a, b, c]
            first = self.const(float(d[1]))
            for v in d[2]:
                self.const(float(v))
            return self.pref(PK_SMD, first), d
        # generic callable: tabulated per run
        fn = d[1]
        slot = self.const(0.0)
        pi = self.pref(PK_TABLE, len(self.table_params))
        self.table_params.append((pi, slot, fn))
        return pi, ("table", fn)

    def add_spline(self, trace) -> int:
        t, c, k = trace.bspline()
        key = (len(t), k, t.tobytes())
        if key not in self.spl_sets:
            self.spl_sets[key] = self._t_len
            self.spl_t.append(t)
            self._t_len += len(t)
        t_off = self.spl_sets[key]
        self.spl_c.append(c)
        self.spl_trace.append((t_off, len(t), self._c_len, k))
        self._c_len += len(c)
        self.spl_fill.append((complex(trace.y[0]), complex(trace.y[-1])))
        self.spl_range.append((float(trace.x[0]), float(trace.x[-1])))
        return len(self.spl_trace) - 1

    # ---- values ----------------------------------------------------------------------------
    def const_value(self, v) -> int:
        self.const_vals.append(complex(v))
        return len(self.const_vals) - 1

    def dyn_values(self, count: int) -> int:
        """Reserve `count` consecutive per-point values; ids are offset by the number of constant
        values when the plan is finalised, so negative placeholders are handed out here."""
        base = self.dyn_count
        self.dyn_count += count
        return base


_DYN = 1 << 28          # marker added to per-point value ids until the constant count is known


class _ValueAtZero:
    """`p.value` (= p(0), numeric.py:116-117) as a per-run scalar: `resolve_run` turns a callable that returns
    a scalar into a constant of that run."""

    def __init__(self, p):
        self._p = p

    def __call__(self, f):
        return np.complex128(self._p.value if hasattr(self._p, "value") else self._p(0))


def describe_foreign(p):
    """Structural description of a parameter object that does not describe itself -- the reference's own
    `heavi.numeric` classes when `Network.run_sp` is bound to this library.  Duck-typed on
    the class name and the attributes those classes have (numeric.py:142-311, lib/optim.py:31-64), so that a
    netlist built with reference objects gets the same device program as one built with heavi_b200's."""
    name = type(p).__name__
    inf_ok = np.isinf(getattr(p, "_inf", np.inf))
    if name == "Scalar" and hasattr(p, "_value"):
        return ("const", complex(p(np.float64(1.0))))
    if name in ("Random", "Param") and hasattr(p, "_value"):
        return ("sample", p, "id") if inf_ok else ("table", p)
    if name == "OptimVar" and hasattr(p, "_value") and getattr(p, "mapping", None) is None:
        return ("sample", p, "id")
    if name in ("Negative", "Inverse") and hasattr(p, "_simparam"):
        inner = p._simparam.describe() if hasattr(p._simparam, "describe") else describe_foreign(p._simparam)
        if inner[0] == "sample" and inner[2] == "id" and inf_ok:
            return ("sample", inner[1], "neg" if name == "Negative" else "inv")
        return ("table", p)
    if name == "Function" and hasattr(p, "_func"):
        fn = p._func
        # Network.transmissionline's closure (network.py:627-633): beta(f) = TWOPI * f / c0 * sqrt(func_er(f))
        if getattr(fn, "__qualname__", "").endswith("transmissionline.<locals>.beta") and getattr(fn, "__closure__", None):
            cells = dict(zip(fn.__code__.co_freevars, (c.cell_contents for c in fn.__closure__)))
            er, c0 = cells.get("func_er"), cells.get("c0")
            if er is not None and c0 == _C0 and inf_ok:
                inner = er.describe() if hasattr(er, "describe") else describe_foreign(er)
                if inner[0] == "const":
                    return ("beta", inner[1])
        return ("table", p)
    if hasattr(p, "bspline") and p.bspline() is not None:
        return ("spline", p)
    if name == "interp1d" and hasattr(p, "_spline") and hasattr(p._spline, "t"):
        # scipy.interpolate.interp1d(kind='cubic', bounds_error=False, fill_value=(y[0], y[-1])): what
        # lib/touchstone.py:258 hands to n_port_S -- a B-spline with constant end fill, evaluated on the device
        tr = _Interp1dTrace(p)
        if tr.ok:
            return ("spline", tr)
    if callable(p):
        return ("table", p)
    return ("const", complex(p))


class _Interp1dTrace:
    """Adapter: a scipy `interp1d` spline seen as the (x, y, bspline()) trace `_Builder.add_spline` reads."""

    def __init__(self, ip):
        self._ip = ip
        self.x = np.asarray(ip.x, dtype=np.float64)
        self.y = np.asarray(ip.y, dtype=np.complex128).reshape(-1)
        spl = ip._spline
        self._tck = (np.asarray(spl.t, dtype=np.float64), np.asarray(spl.c, dtype=np.complex128).reshape(len(spl.c)), int(spl.k))
        fv = getattr(ip, "fill_value", None)
        try:
            lo, hi = fv
            same = complex(np.asarray(lo).reshape(-1)[0]) == complex(self.y[0]) and complex(np.asarray(hi).reshape(-1)[0]) == complex(self.y[-1])
        except Exception:
            same = False
        self.ok = bool(same and not getattr(ip, "bounds_error", True) and self.y.shape == self.x.shape and self._tck[2] <= 3)

    def __call__(self, f):
        return self._ip(f)

    def bspline(self):
        return self._tck


def _is_static(desc) -> bool:
    return desc[0] == "const"


def _component_values(b: _Builder, kind: str, par: dict, nterm: int):
    """Emit the value program of one component.  Returns a (k,k) array of value ids (or None for a
    structural zero) and a matching array of negate flags, k = len(ids)."""
    f1 = np.array([1.0])                     # dummy frequency for plan-time evaluation of constants

    def block2(vid):
        ids = np.array([[vid, vid], [vid, vid]], dtype=np.int64)
        neg = np.array([[0, 1], [1, 0]], dtype=np.int8)
        return ids, neg

    if kind == "port":
        pi, d = b.param(par["Z0"])
        if _is_static(d):
            g = b.const_value(1 / np.float64(np.real(d[1])) if np.imag(d[1]) == 0 else 1 / d[1])
        else:
            g = _DYN + b.dyn_values(1)
            b.ops.append([OP_RECIP, g - _DYN, pi, 0, 0, 0, 0, 0])
        one = 0
        ids = np.full((4, 4), -1, dtype=np.int64)
        neg = np.zeros((4, 4), dtype=np.int8)
        # conductance between sig(0) and int(1); ideal source int(1)->gnd(2) on row/col src(3)
        for (r, c, v, s) in [(0, 0, g, 0), (0, 1, g, 1), (1, 0, g, 1), (1, 1, g, 0),
                             (1, 3, one, 0), (2, 3, one, 1), (3, 1, one, 0), (3, 2, one, 1)]:
            ids[r, c], neg[r, c] = v, s
        return ids, neg, pi

    if kind in ("resistor", "admittance", "impedance", "capacitor", "inductor"):
        key = {"resistor": "G", "admittance": "Y", "impedance": "Z", "capacitor": "C", "inductor": "L"}[kind]
        p = par[key]
        if kind == "resistor":
            # R is a plain number in the reference (App. A), so G(f) is the same scalar at every f
            return (*block2(b.const_value(complex(np.asarray(p(f1)).reshape(-1)[0]))), None)
        pi, d = b.param(p)
        if _is_static(d):
            v = d[1]
            v = np.float64(v.real) if v.imag == 0 else np.complex128(v)
            # frequency-independent kinds fold to a constant; C and L always depend on f
            if kind == "admittance":
                return (*block2(b.const_value(complex(v))), None)
            if kind == "impedance":
                return (*block2(b.const_value(complex(1 / v))), None)
        opcode = {"admittance": OP_COPY, "impedance": OP_RECIP, "capacitor": OP_CAP, "inductor": OP_IND}[kind]
        vid = _DYN + b.dyn_values(1)
        b.ops.append([opcode, vid - _DYN, pi, 0, 0, 0, 0, 0])
        return (*block2(vid), None)

    if kind == "tline":
        pz, _ = b.param(par["Z0"])
        pb, _ = b.param(par["beta"])
        # the reference reads `self.length.value` (= length(0)) inside every run_sp (base/tl.py:52): a Param /
        # Random / optimiser variable as length is a sample column, a callable is evaluated at f = 0 per run
        length = par["length"]
        if callable(length) and b.describe(length)[0] not in ("const", "sample"):
            length = _ValueAtZero(length)
        plen, _ = b.param(length)
        base = b.dyn_values(9)
        b.ops.append([OP_TLINE, base, pz, pb, plen, 0, 0, 0])
        v = _DYN + base
        # [y11 y12 -(y11+y12)] [y21 y22 -(y21+y22)] [-(y11+y21) -(y12+y22) corner]
        ids = np.array([[v + 0, v + 1, v + 4], [v + 2, v + 3, v + 5], [v + 6, v + 7, v + 8]], dtype=np.int64)
        return ids, np.zeros((3, 3), dtype=np.int8), None

    if kind in ("nport_s", "nport_y"):
        funcs = par["S"] if kind == "nport_s" else par["Y"]
        Pn = len(funcs)
        if Pn > MAX_NPORT:
            raise ValueError(f"N-port blocks with more than {MAX_NPORT} ports are not supported on the device")
        first = len(b.prefs)
        for i in range(Pn):
            for j in range(Pn):
                b.param(funcs[i][j])
        # every b.param() call appended exactly one pref -> they are consecutive, row-major
        assert len(b.prefs) == first + Pn * Pn
        base = b.dyn_values((Pn + 1) * (Pn + 1))
        if kind == "nport_s":
            z0 = b.const(complex(par["Z0"]))
            b.coops.append([COOP_NPORT_S, base, Pn, first, z0, 0, 0, 0])
            b.coop_scratch = max(b.coop_scratch, 2 * Pn * Pn)
        else:
            b.coops.append([COOP_NPORT_Y, base, Pn, first, 0, 0, 0, 0])
            b.coop_scratch = max(b.coop_scratch, Pn * Pn)
        ids = (_DYN + base + np.arange((Pn + 1) * (Pn + 1), dtype=np.int64)).reshape(Pn + 1, Pn + 1)
        return ids, np.zeros((Pn + 1, Pn + 1), dtype=np.int8), None

    raise NotImplementedError(f"component kind {kind!r} has no S-parameter stamp")


def norton_legal(table: StampTable) -> bool:
    """A port can be folded into a Norton source when its internal node is private to it."""
    if table.P == 0:
        return False
    used = {}
    for ci, ids in enumerate(table.ids):
        for i in set(ids):
            used.setdefault(i, set()).add(ci)
    port_comp = [ci for ci, k in enumerate(table.kinds) if k == "port"]
    seen_int = set()
    for ci in port_comp:
        sig, internal, gnd, src = table.ids[ci]
        if internal == 0 or internal in (sig, gnd) or internal in seen_int:
            return False
        if used[internal] != {ci} or used[src] != {ci}:
            return False
        seen_int.add(internal)
    return True


BAND_MIN_ROWS = 33        # systems the register-resident kernels cannot hold
BAND_SMALL_MIN_ROWS = 9   # from here on the banded kernel beats the register kernels on large batches
BAND_MAX_HALF_WIDTH = 8   # widest band the banded kernel is instantiated for


def _band_order(table: StampTable, rowmap: np.ndarray, mode: str, small: bool = False):
    """Reverse Cuthill-McKee ordering of the system rows (a symmetric permutation of the unknowns,
    which leaves S unchanged).  Returns (new rowmap, half bandwidth) or None when the structure
    is not narrow-banded.  Chains -- ladders, cascaded lines, filters -- come out tri-/penta-diagonal."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    n = int(rowmap.max()) + 1
    rr, cc = [], []
    for kind, ids in zip(table.kinds, table.ids):
        ids = [int(i) for i in ids]
        if kind == "port" and mode == "norton":
            ids = [ids[0], ids[2]]                      # only sig and the reference node stay coupled
        rows = [int(rowmap[i]) for i in ids if rowmap[i] >= 0]
        for a in rows:
            for c in rows:
                rr.append(a); cc.append(c)
    if not rr:
        return None
    A = coo_matrix((np.ones(len(rr), dtype=np.int8), (rr, cc)), shape=(n, n)).tocsr()
    coo = A.tocoo()
    perm = np.asarray(reverse_cuthill_mckee(A, symmetric_mode=True), dtype=np.int64)   # new -> old
    pos = np.empty(n, dtype=np.int64)
    pos[perm] = np.arange(n)                            # old -> new
    band = int(np.abs(pos[coo.row] - pos[coo.col]).max())
    natural = int(np.abs(coo.row - coo.col).max())
    if natural <= band:                                 # the netlist was already written as a chain
        pos, band = np.arange(n, dtype=np.int64), natural
    # a narrow band, not a small dense block; `small` (order="rcm" asked for explicitly) only needs the band
    # to be narrower than the matrix
    if band > BAND_MAX_HALF_WIDTH or (2 * band + 1 > n if small else 2 * (3 * band + 1) > n):
        return None
    new = rowmap.copy()
    live = rowmap >= 0
    new[live] = pos[rowmap[live]]
    return new, band


def compile_plan(table: StampTable, mode: str = "auto", pad_to: tuple[int, int] | None = None,
                 order: str = "auto") -> DevicePlan:
    """`pad_to=(rows, rhs_columns)` lays the system out in a larger fixed shape: extra unknowns are
    identity rows (x = 0) and extra right-hand sides are zero columns, so the solution of the real
    unknowns is unchanged while the register-resident kernel runs without bounds checks.

    `order`: "natural" keeps the reference's node numbering for the system rows; "rcm" applies a
    bandwidth-reducing ordering when the netlist is narrow-banded (what the banded kernel needs);
    "auto" does so for systems too large for the register-resident kernels."""
    if mode == "auto":
        mode = "norton" if norton_legal(table) else "full"
    if mode == "norton" and not norton_legal(table):
        raise ValueError("Norton port folding is not legal for this netlist")
    if mode not in ("full", "norton", "dump"):
        raise ValueError(mode)
    b = _Builder()
    NM = table.N + table.M
    P = table.P

    # ---- MNA index -> system row -----------------------------------------------------------
    rowmap = np.full(NM, ROW_GROUND, dtype=np.int64)
    if mode == "dump":
        rowmap[:] = np.arange(NM)
    elif mode == "full":
        rowmap[1:] = np.arange(NM - 1)
    else:
        dropped = set(int(s) for s in table.port_src) | set(int(r[2]) for r in table.port_table)
        keep = [i for i in range(1, NM) if i not in dropped]
        rowmap[keep] = np.arange(len(keep))
    n_real = int(rowmap.max()) + 1 if NM > 0 else 0
    if order not in ("auto", "natural", "rcm"):
        raise ValueError(order)
    band = -1
    if mode != "dump" and pad_to is None and (order == "rcm" or (order == "auto" and n_real >= BAND_MIN_ROWS)):
        # N-port blocks: the banded kernel evaluates two-ports per thread; larger blocks need the lane-cooperative path
        if all(len(ids) == 3 for k, ids in zip(table.kinds, table.ids) if k in ("nport_s", "nport_y")):
            res = _band_order(table, rowmap, mode, small=(order == "rcm"))
            if res is not None:
                rowmap, band = res
    n, n_rhs = n_real, (0 if mode == "dump" else P)
    if pad_to is not None and mode != "dump":
        if pad_to[0] < n_real or pad_to[1] < P:
            raise ValueError(f"pad_to={pad_to} is smaller than the system ({n_real}, {P})")
        n, n_rhs = int(pad_to[0]), int(pad_to[1])
    ncols = n + n_rhs

    cells: dict[tuple[int, int], list[tuple[int, int]]] = {}

    def add(r, c, vid, neg):
        cells.setdefault((int(r), int(c)), []).append((int(vid), int(neg)))

    port_pref = {}
    for ci, (kind, ids, par) in enumerate(zip(table.kinds, table.ids, table.params)):
        vids, negs, zpref = _component_values(b, kind, par, len(ids))
        if kind == "port":
            port_pref[par["port"]] = (zpref, vids[0, 0], ids)
            if mode == "norton":
                sig, internal, gnd, src = ids
                g = vids[0, 0]
                # conductance g between sig and the port's reference node
                pairs = [(sig, sig, 0), (sig, gnd, 1), (gnd, sig, 1), (gnd, gnd, 0)]
                last = {}
                for (mr, mc, s) in pairs:
                    last[(mr, mc)] = s            # numpy-style last write wins if sig == gnd
                for (mr, mc), s in last.items():
                    if rowmap[mr] >= 0 and rowmap[mc] >= 0:
                        add(rowmap[mr], rowmap[mc], g, s)
                continue
        # last (row-major) block position writing each distinct cell wins (numpy `+=` semantics)
        k = len(ids)
        last = {}
        for a in range(k):
            for c in range(k):
                last[(ids[a], ids[c])] = (a, c)
        for (mr, mc), (a, c) in last.items():
            if vids[a, c] < 0:
                continue
            if rowmap[mr] < 0 or rowmap[mc] < 0:
                continue
            add(rowmap[mr], rowmap[mc], vids[a, c], negs[a, c])

    # ---- right-hand sides and S epilogue tables --------------------------------------------
    port_rows = np.full((P, 3), ROW_GROUND, dtype=np.int32)
    port_zpref = np.zeros(P, dtype=np.int32)
    for q in range(P):
        iport = int(table.port_table[q, 0]) + 1
        zpref, g, ids = port_pref[iport]
        sig, internal, gnd, src = ids
        port_zpref[q] = zpref
        port_rows[q, 0] = rowmap[sig]
        port_rows[q, 2] = rowmap[gnd]
        if mode == "norton":
            port_rows[q, 1] = ROW_VIRTUAL_INT
            if rowmap[sig] >= 0:
                add(rowmap[sig], n + q, g, 0)
            if rowmap[gnd] >= 0:
                add(rowmap[gnd], n + q, g, 1)
        else:
            port_rows[q, 1] = rowmap[internal]
            if mode == "full":
                add(rowmap[src], n + q, 0, 0)          # unit excitation on the port's source row

    for r in range(n_real, n):                       # identity padding rows (value id 0 == 1+0j)
        add(r, r, 0, 0)

    # ---- finalise value ids and CSR ----------------------------------------------------------
    nvc = len(b.const_vals)

    def final(vid):
        return vid - _DYN + nvc if vid >= _DYN else vid

    # cells sorted by (row, column): one pass builds the cell CSR and the per-row lists (the matrix is ~1 %
    # dense for large netlists, so nothing here loops over the n x ncols grid)
    items = sorted(cells.items())
    counts = np.zeros(n * ncols, dtype=np.int64)
    ent, rent = [], []
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    for (r, c), lst in items:
        counts[r * ncols + c] = len(lst)
        if c >= (1 << 11):
            raise ValueError("plan too large for the packed row-entry format")
        for k, (vid, neg) in enumerate(lst):
            v = final(vid)
            if v >= (1 << 19):
                raise ValueError("plan too large for the packed row-entry format")
            ent.append((v << 1) | neg)
            # per-row lists: the same entries with the column and an end-of-cell flag packed in
            word = (neg & 1) | (v << 1) | (c << 20) | ((1 << 31) if k == len(lst) - 1 else 0)
            rent.append(word - (1 << 32) if word >= (1 << 31) else word)
        row_ptr[r + 1] += len(lst)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    cell_ptr = np.zeros(n * ncols + 1, dtype=np.int32)
    np.cumsum(counts, out=cell_ptr[1:])
    ops = np.array(b.ops, dtype=np.int32).reshape(len(b.ops), OP_WORDS)
    coops = np.array(b.coops, dtype=np.int32).reshape(len(b.coops), COOP_WORDS)
    if len(ops):
        ops[:, 1] += nvc
    if len(coops):
        coops[:, 1] += nvc

    # ---- fast value ops: split off the simple ops with a real-constant or plain-sample argument ----
    consts_arr = np.array(b.consts, dtype=np.complex128).reshape(len(b.consts))
    fcode, fout, farg, slow = [], [], [], []
    for op in ops:
        code, out, pi = int(op[0]), int(op[1]), int(op[2])
        kind, idx = b.prefs[pi] if code in (OP_COPY, OP_RECIP, OP_CAP, OP_IND) else (None, None)
        if kind == PK_CONST and consts_arr[idx].imag == 0.0:
            fcode.append(code); fout.append(out - nvc); farg.append(float(consts_arr[idx].real))
        elif kind == PK_SAMPLE:
            fcode.append(code | 4); fout.append(out - nvc); farg.append(float(idx))
        else:
            slow.append(op)
    ops = np.array(slow, dtype=np.int32).reshape(len(slow), OP_WORDS)

    # ---- closed-form epilogue ------------------------------------------------------------------
    epi_simple, epi_port_of_row, epi_scale = 0, np.full(max(n, 1), -1, dtype=np.int32), np.ones((max(P, 1), max(P, 1)))
    if mode == "norton" and P > 0:
        z = []
        ok = True
        for q in range(P):
            kind, idx = b.prefs[port_zpref[q]]
            if kind != PK_CONST or consts_arr[idx].imag != 0.0 or not (consts_arr[idx].real > 0.0):
                ok = False
                break
            z.append(float(consts_arr[idx].real))
            if port_rows[q, 2] != ROW_GROUND or port_rows[q, 0] < 0 or epi_port_of_row[port_rows[q, 0]] != -1:
                ok = False
                break
            epi_port_of_row[port_rows[q, 0]] = q
        if ok:
            epi_simple = 1
            zz = np.array(z)
            epi_scale = np.sqrt(np.abs(zz))[None, :] / np.sqrt(np.abs(zz))[:, None]      # [j, i]
        else:
            epi_port_of_row[:] = -1

    nT = len(b.spl_trace)
    return DevicePlan(
        mode=mode, n=n, ncols=ncols, P=P, NM=NM,
        consts=np.array(b.consts, dtype=np.complex128).reshape(len(b.consts)),
        prefs=np.array(b.prefs, dtype=np.int32).reshape(len(b.prefs), 2),
        const_vals=np.array(b.const_vals, dtype=np.complex128),
        n_dyn=b.dyn_count, ops=ops, coops=coops,
        cell_ptr=cell_ptr, cell_ent=np.array(ent, dtype=np.int32),
        row_ptr=row_ptr, row_ent=np.array(rent, dtype=np.int32),
        port_rows=port_rows, port_zpref=port_zpref,
        spl_t=np.concatenate(b.spl_t) if b.spl_t else np.zeros(0),
        spl_c=np.concatenate(b.spl_c) if b.spl_c else np.zeros(0, dtype=np.complex128),
        spl_trace=np.array(b.spl_trace, dtype=np.int32).reshape(nT, 4),
        spl_fill=np.array(b.spl_fill, dtype=np.complex128).reshape(nT, 2),
        spl_range=np.array(b.spl_range, dtype=np.float64).reshape(nT, 2),
        sample_owners=b.sample_owners, table_params=b.table_params,
        n_static_consts=len(b.consts), coop_scratch=b.coop_scratch,
        fop_code=np.array(fcode, dtype=np.int32), fop_out=np.array(fout, dtype=np.int32),
        fop_arg=np.array(farg, dtype=np.float64),
        epi_simple=epi_simple, epi_port_of_row=epi_port_of_row, epi_scale=np.ascontiguousarray(epi_scale, dtype=np.float64),
        n_real=n_real, max_nport=max([int(c[2]) for c in b.coops], default=0), band=band,
    )


# ==================================================================================================
# per-run parameter resolution (host side of a run)
# ==================================================================================================
def resolve_run(plan: DevicePlan, f: np.ndarray, drivers=None, values=None):
    """Evaluate what only the host can evaluate for this run.

    Returns (consts, prefs, tables[nT_used, nF], samples[nS, nR]).  Callable parameters are called
    once on the whole float64 frequency array -- the reference's own arithmetic, so tabulated
    values are bit-identical; a callable that returns a scalar becomes a constant for this run.
    """
    consts = plan.consts.copy()
    prefs = plan.prefs.copy()
    rows = []
    seen = {}                                  # a callable shared by several components is evaluated (and uploaded) once
    for (pi, slot, fn) in plan.table_params:
        # `Function` wrappers of one user callable (each component wraps its own) count as the same table
        inner = getattr(fn, "_func", None)
        key = (id(inner), getattr(fn, "_inf", None)) if inner is not None and type(fn).__name__ == "Function" else id(fn)
        if key not in seen:
            v = np.asarray(fn(f))
            if v.ndim == 0 or v.size == 1 and f.size != 1:
                seen[key] = (None, complex(v.reshape(-1)[0]))
            else:
                if v.shape != f.shape:
                    v = np.broadcast_to(v, f.shape)
                seen[key] = (len(rows), None)
                rows.append(np.ascontiguousarray(v, dtype=np.complex128))
        row, scalar = seen[key]
        if row is None:
            consts[slot] = scalar
            prefs[pi] = (PK_CONST, slot)
        else:
            prefs[pi] = (PK_TABLE, row)
    tables = np.stack(rows) if rows else np.zeros((0, f.shape[0]), dtype=np.complex128)

    nR = len(plan.sample_owners)
    if values is None:
        nS = 1
        samples = np.empty((1, nR), dtype=np.float64)
        col_of = {}
    else:
        values = np.asarray(values, dtype=np.float64)
        nS = values.shape[0]
        samples = np.empty((nS, nR), dtype=np.float64)
        col_of = {id(d): k for k, d in enumerate(drivers)}
    for c, owner in enumerate(plan.sample_owners):
        if id(owner) in col_of:
            samples[:, c] = values[:, col_of[id(owner)]]
        else:
            v = owner._value
            if not isinstance(v, (int, float, np.integer, np.floating)):
                if isinstance(v, complex) or isinstance(v, np.complexfloating):
                    raise TypeError("complex Monte-Carlo / sweep values are not supported on the device")
                raise ValueError("Uninitialized value")
            samples[:, c] = float(v)
    return consts, prefs, tables, samples


# ==================================================================================================
# stamp table with tagged parameters: what the C-ABI's own compiler takes (hvb_plan_create_from_stamps)
# ==================================================================================================
COMP_KINDS = {"port": 0, "resistor": 1, "admittance": 2, "impedance": 3, "capacitor": 4, "inductor": 5, "tline": 6,
              "nport_s": 7, "nport_y": 8}
PAR_DTYPE = np.dtype([("kind", np.int32), ("index", np.int32), ("v", np.float64, (4,))])


@dataclass
class TaggedTable:
    comp_kind: np.ndarray
    comp_port: np.ndarray
    ids_off: np.ndarray
    ids: np.ndarray
    par_off: np.ndarray
    par: np.ndarray               # PAR_DTYPE
    port_table: np.ndarray
    port_src: np.ndarray
    N: int
    M: int
    spl_t: np.ndarray
    spl_c: np.ndarray
    spl_trace: np.ndarray
    spl_fill: np.ndarray
    spl_range: np.ndarray
    sample_owners: list
    table_fns: list


def tag_table(table: StampTable) -> TaggedTable:
    """Turn every parameter object of a stamp table into a tagged slot (CONST / SAMPLE / TABLE / SPLINE / BETA / SMD),
    in the order the compiler consumes them.  Host-side bookkeeping stays here: which object owns which sample
    column, which callable fills which table row, the B-spline data of the traces."""
    b = _Builder()                          # only its describe() / add_spline() are used
    f1 = np.array([1.0])
    pars, par_off, kinds, comp_port = [], [0], [], []
    sample_owners, table_fns = [], []

    def tag(p):
        d = b.describe(p)
        t = d[0]
        if t == "const":
            v = complex(d[1])
            return (PK_CONST, 0, (v.real, v.imag, 0.0, 0.0))
        if t == "sample":
            owner, op = d[1], d[2]
            for col, o in enumerate(sample_owners):
                if o is owner:
                    break
            else:
                sample_owners.append(owner)
                col = len(sample_owners) - 1
            return ({"id": PK_SAMPLE, "neg": PK_SAMPLE_NEG, "inv": PK_SAMPLE_INV}[op], col, (0.0, 0.0, 0.0, 0.0))
        if t == "beta":
            er = complex(d[1])
            with np.errstate(invalid="ignore"):
                root = complex(np.sqrt(np.float64(er.real)) if er.imag == 0 else np.sqrt(np.complex128(er)))
            return (PK_BETA, 0, (root.real, root.imag, 0.0, 0.0))
        if t == "spline":
            return (PK_SPLINE, b.add_spline(d[1]), (0.0, 0.0, 0.0, 0.0))
        if t == "beta_mul":
            return (PK_BETA_MUL, 0, (float(d[1]), 0.0, 0.0, 0.0))
        if t == "smd":
            a, bb, c = (float(x) for x in d[2])
            return (PK_SMD, int(d[1]), (a, bb, c, 0.0))
        table_fns.append(d[1])
        return (PK_TABLE, len(table_fns) - 1, (0.0, 0.0, 0.0, 0.0))

    for kind, par in zip(table.kinds, table.params):
        kinds.append(COMP_KINDS[kind])
        comp_port.append(int(par.get("port", 0)) if kind == "port" else 0)
        if kind == "port":
            pars.append(tag(par["Z0"]))
        elif kind == "resistor":
            g = complex(np.asarray(par["G"](f1)).reshape(-1)[0])
            pars.append((PK_CONST, 0, (g.real, g.imag, 0.0, 0.0)))
        elif kind in ("admittance", "impedance", "capacitor", "inductor"):
            pars.append(tag(par[{"admittance": "Y", "impedance": "Z", "capacitor": "C", "inductor": "L"}[kind]]))
        elif kind == "tline":
            length = par["length"]
            if callable(length) and b.describe(length)[0] not in ("const", "sample"):
                length = _ValueAtZero(length)
            pars.extend([tag(par["Z0"]), tag(par["beta"]), tag(length)])
        elif kind in ("nport_s", "nport_y"):
            funcs = par["S"] if kind == "nport_s" else par["Y"]
            for row in funcs:
                for fn in row:
                    pars.append(tag(fn))
            if kind == "nport_s":
                z0 = complex(par["Z0"])
                pars.append((PK_CONST, 0, (z0.real, z0.imag, 0.0, 0.0)))
        else:
            raise NotImplementedError(f"component kind {kind!r} has no S-parameter stamp")
        par_off.append(len(pars))
    par_arr = np.zeros(max(len(pars), 1), dtype=PAR_DTYPE)
    for i, (k, idx, v) in enumerate(pars):
        par_arr[i] = (k, idx, v)
    off = np.cumsum([0] + [len(i) for i in table.ids]).astype(np.int32)
    nT = len(b.spl_trace)
    return TaggedTable(
        comp_kind=np.array(kinds, dtype=np.int32), comp_port=np.array(comp_port, dtype=np.int32), ids_off=off,
        ids=np.array([i for lst in table.ids for i in lst], dtype=np.int32), par_off=np.array(par_off, dtype=np.int32), par=par_arr,
        port_table=np.ascontiguousarray(table.port_table, dtype=np.int32), port_src=np.ascontiguousarray(table.port_src, dtype=np.int32),
        N=int(table.N), M=int(table.M),
        spl_t=np.concatenate(b.spl_t) if b.spl_t else np.zeros(0),
        spl_c=np.concatenate(b.spl_c) if b.spl_c else np.zeros(0, dtype=np.complex128),
        spl_trace=np.array(b.spl_trace, dtype=np.int32).reshape(nT, 4),
        spl_fill=np.array(b.spl_fill, dtype=np.complex128).reshape(nT, 2),
        spl_range=np.array(b.spl_range, dtype=np.float64).reshape(nT, 2),
        sample_owners=sample_owners, table_fns=table_fns)
