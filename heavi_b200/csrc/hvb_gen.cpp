// hvb_gen.cpp -- kernel D source generator (see hvb_gen.h).
//
// Algorithm the generated kernel runs, one THREAD per (sample, frequency) point:
//
//   The plan compiler's bandwidth-reducing order makes ladders, filters and cascaded lines
//   tridiagonal.  Row-oriented elimination with partial pivoting (the zgttrf recurrence: the pivot of
//   column k is the larger of row k and row k+1, U gets a second super-diagonal) walks the chain once,
//   front to back, holding two rows in registers.  What the S epilogue needs are the voltages of the
//   P port nodes only (solvers/spfn.py:159-176), x_o = e_o^T U^-1 y.  Instead of storing every pivot
//   row for a back substitution, each port row o carries w_o = e_o^T U^-1 forward (a forward
//   substitution against the columns of U: w_o[k] = (d_ok - w_o[k-1] U[k-1][k] - w_o[k-2] U[k-2][k]) / U[k][k])
//   and accumulates x_o += w_o[k] y_k on the fly, so nothing but S ever reaches HBM.
//
//   Right-hand sides: a port's Norton source enters at its own row.  Once a column's entry row has
//   passed, its entries in the two live rows are (c_j, 0); one elimination step multiplies every such
//   column by the same factor (-m without a row exchange, 1 with), so all "old" columns stay
//   proportional: c_j = gamma_j * z with one scalar z per point.  Per step the kernel updates z and
//   T_o += w_o[k] * zeta_k; only when a new port joins (P times per point) it folds
//   acc[o][j] += gamma_j * T_o.  Work per row is O(active ports), not O(ports^2).
//
//   Two-sided form (long chains, rolled): a port's recurrence w_o runs from its row to the END of the chain,
//   so the work per row grows with the ports already passed.  The chain is therefore cut between two
//   non-port rows m | m+1; the front half (rows 0..m) is walked forwards and the back half (rows n-1..m+1)
//   backwards, each by the one-sided algorithm above with ONE extra unit source at its last row.  That
//   extra tap yields, per half H with A_H its own tridiagonal block, delta_H = (A_H^-1)[last][last],
//   beta_H,j = (A_H^-1 b_j)[last] and alpha_H,o = (A_H^-1)[o][last]; with v = A[m][m+1], u = A[m+1][m] the
//   interface voltages follow from a 2 x 2 system,
//       xi_j = x_m,j = (beta_F,j - v delta_F beta_B,j) / (1 - v u delta_F delta_B),  eta_j = x_m+1,j = beta_B,j - u delta_B xi_j,
//   and the port voltages are x_o,j = X_H[o][j] - alpha_H,o * (v eta_j | u xi_j).  Every recurrence stops at
//   the cut: about half the border work and half the live registers.  The cut restricts pivoting (rows m and
//   m+1 are never exchanged); the kernel checks the cancellation in the 2 x 2 solve and in every x_o,j and
//   sets HVB_STATUS_RERUN_BIT beyond 10^3 -- the library then repeats the call with the one-sided kernel.
//
// Same pivots as a dense LU with partial pivoting of the same row order (the rest of column k is
// structurally zero), same stamp formulas and summation order as the reference's closures
// (base/*.py); transmission lines use the single-reciprocal form (eval_tline_fast).
#include "hvb_gen.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "plan_format.h"

#ifndef HVB_GEN_UNROLL_MAX
#define HVB_GEN_UNROLL_MAX 40
#endif
#define HVB_GEN_MAX_PORTS 16
#define HVB_GEN_REG_ACC_PORTS 4
#define HVB_GEN_VID_ONE 0x7FFFF        // value id of the literal 1 (unit source of a half chain's extra tap)

void GenPlan::from_desc(const hvb_plan_desc& d) {
    n = d.n; ncols = d.ncols; P = d.P; nvc = d.n_const_vals; ndyn = d.n_dyn; nops = d.n_ops; nfast = d.n_fast_ops;
    ncoops = d.n_coops; epi_simple = d.epi_simple; nprefs = d.n_prefs;
    const_vals.assign(d.const_vals, d.const_vals + 2 * (size_t)nvc);
    ops.assign(d.ops, d.ops + (size_t)nops * HVB_OP_WORDS);
    row_ptr.assign(d.row_ptr, d.row_ptr + n + 1);
    row_ent.assign(d.row_ent, d.row_ent + d.n_row_ent);
    fop_code.assign(d.fop_code, d.fop_code + nfast);
    fop_out.assign(d.fop_out, d.fop_out + nfast);
    fop_arg.assign(d.fop_arg, d.fop_arg + nfast);
    port_rows.assign(d.port_rows, d.port_rows + 3 * (size_t)P);
    epi_port_of_row.assign(d.epi_port_of_row, d.epi_port_of_row + n);
    epi_scale.assign(d.epi_scale, d.epi_scale + (size_t)std::max(P * P, 1));
    coops.assign(d.coops, d.coops + (size_t)ncoops * HVB_COOP_WORDS);
    spl_trace.assign(d.spl_trace, d.spl_trace + (size_t)d.n_traces * 4);
}

namespace {

struct Entry { int col, vid; bool neg; };
struct Prod { int kind = -1, idx = 0, sub = 0; };      // kind 0: fast op, 1: simple op (sub = output element)

struct Analysis {
    std::vector<std::vector<Entry>> rows;
    std::vector<int> tap_row, tap_port, tap_of_row;
    std::vector<Prod> prod;
    std::vector<int> fop_first, fop_last, op_first, op_last;
    std::vector<unsigned> op_submask;             // outputs of a (line) op that some matrix cell sums
    std::string why;
    bool ok = false;
};

std::string fmt(const char* f, ...) __attribute__((format(printf, 1, 2)));
std::string fmt(const char* f, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return buf;
}

Analysis analyse(const GenPlan& g) {
    Analysis a;
    const int n = g.n, P = g.P;
    auto no = [&](const std::string& w) { a.why = w; return a; };
    if (!g.epi_simple) return no("S epilogue is not the closed form (ports not ground-referenced with constant real Z0, or not a Norton plan)");
    if (g.ncoops) return no("plan holds N-port blocks");
    if (g.ncols - n != P || P < 1) return no("plan is padded or has no ports");
    if (P > HVB_GEN_MAX_PORTS) return no("too many ports");
    if (n < 2) return no("fewer than two unknowns");
    a.rows.resize(n);
    for (int r = 0; r < n; ++r)
        for (int e = g.row_ptr[r]; e < g.row_ptr[r + 1]; ++e) {
            const uint32_t w = (uint32_t)g.row_ent[e];
            Entry en{(int)((w >> 20) & 0x7FFu), (int)((w >> 1) & 0x7FFFFu), (w & 1u) != 0};
            if (en.col < n && std::abs(en.col - r) > 1) return no("matrix is not tridiagonal in the plan's row order");
            a.rows[r].push_back(en);
        }
    // ports: one Norton source entry per right-hand-side column, on the port's own signal row
    std::vector<int> rhs_count(P, 0);
    for (int r = 0; r < n; ++r)
        for (const Entry& en : a.rows[r])
            if (en.col >= n && en.col != g.cut_col) {
                const int q = en.col - n;
                if (q >= P || g.port_rows[3 * q] != r || en.neg) return no("right-hand side is not one source entry per port row");
                rhs_count[q]++;
            }
    a.tap_of_row.assign(n, -1);
    std::vector<std::pair<int, int>> taps;
    for (int q = 0; q < P; ++q) {
        const int r = g.port_rows[3 * q];
        if (r < 0 || r >= n || rhs_count[q] < 1 || g.epi_port_of_row[r] != q) return no("port without a private signal row");
        taps.push_back({r, q});
    }
    std::sort(taps.begin(), taps.end());
    for (size_t f = 0; f < taps.size(); ++f) {
        if (f && taps[f].first == taps[f - 1].first) return no("two ports on one row");
        a.tap_row.push_back(taps[f].first);
        a.tap_port.push_back(taps[f].second);
        a.tap_of_row[taps[f].first] = (int)f;
    }
    // value producers and their live ranges (rows)
    a.prod.assign(std::max(g.ndyn, 1), Prod{});
    for (int o = 0; o < g.nfast; ++o) {
        const int v = g.fop_out[o];
        if (v < 0 || v >= g.ndyn) return no("bad fast-op output");
        a.prod[v] = Prod{0, o, 0};
    }
    for (int o = 0; o < g.nops; ++o) {
        const int32_t* op = &g.ops[(size_t)o * HVB_OP_WORDS];
        const int out = op[1] - g.nvc, cnt = op[0] == OP_TLINE ? 9 : 1;
        for (int e = 0; e < cnt; ++e) {
            if (out + e < 0 || out + e >= g.ndyn) return no("bad op output");
            a.prod[out + e] = Prod{1, o, e};
        }
    }
    a.fop_first.assign(std::max(g.nfast, 1), n); a.fop_last.assign(std::max(g.nfast, 1), -1);
    a.op_first.assign(std::max(g.nops, 1), n); a.op_last.assign(std::max(g.nops, 1), -1);
    a.op_submask.assign(std::max(g.nops, 1), 0u);
    for (int r = 0; r < n; ++r)
        for (const Entry& en : a.rows[r]) {
            if (en.vid < g.nvc || en.vid == HVB_GEN_VID_ONE) continue;
            const Prod& p = a.prod[en.vid - g.nvc];
            if (p.kind < 0) return no("value without a producer");
            std::vector<int>& fi = p.kind == 0 ? a.fop_first : a.op_first;
            std::vector<int>& la = p.kind == 0 ? a.fop_last : a.op_last;
            fi[p.idx] = std::min(fi[p.idx], r);
            la[p.idx] = std::max(la[p.idx], r);
            if (p.kind == 1) a.op_submask[p.idx] |= 1u << p.sub;
        }
    for (int o = 0; o < g.nfast; ++o) if (a.fop_last[o] - a.fop_first[o] > 1) return no("value used in non-adjacent rows");
    for (int o = 0; o < g.nops; ++o) if (a.op_last[o] >= 0 && a.op_last[o] - a.op_first[o] > 1) return no("value used in non-adjacent rows");
    a.ok = true;
    return a;
}

// constants of one row: literals (unrolled) or per-row table slots (rolled)
struct Sink {
    bool rolled;
    std::vector<int32_t>* itab;
    std::vector<double>* dtab;
    int ni = 0, nd = 0;                 // slots used by the current row
    std::string d(double v) {
        if (!rolled) {
            if (std::isnan(v)) return "__longlong_as_double(0x7ff8000000000000LL)";
            if (std::isinf(v)) return v > 0 ? "__longlong_as_double(0x7ff0000000000000LL)" : "__longlong_as_double(0xfff0000000000000LL)";
            return fmt("(%a)", v);
        }
        dtab->push_back(v);
        return fmt("dt[%d]", nd++);
    }
    std::string i(int v) {
        if (!rolled) return fmt("%d", v);
        itab->push_back(v);
        return fmt("it[%d]", ni++);
    }
};

struct ValRef { std::string re, im; };   // "0" = exactly zero, known at generation time

struct Gen {
    const GenPlan& g;
    const Analysis& a;
    const int32_t* prefs;
    bool rolled, scratch;
    int max_vs[2] = {0, 0}, max_vc[2] = {0, 0}, max_vt[2] = {0, 0}, max_vl[2] = {0, 0};
    std::map<int, std::string> fop_name, op_name;     // producer index -> variable (current rows only)
    std::map<int, bool> op_chain_form;                // line op evaluated in the two-value form (y11 = y22, y12 = y21)
    std::map<int, bool> op_imag_only;                 // ... of a lossless line with real Z0: both values are purely imaginary
    long long flops = 0;
    // two-sided form: this Gen serves one half chain
    bool half = false;
    std::vector<int> gport;               // half: local port -> port of the whole plan (-1: the extra tap)
    int Pglob = 0, stash = 0;             // half: ports of the whole plan; a port of the OTHER half (its column is free here)
    int tt_base = 0, gm_base = 0, x_base = 0;   // shared-memory slots of T_o, gamma_j and (half) the extra tap's row

    Gen(const GenPlan& g_, const Analysis& a_, const int32_t* p, bool r, bool s) : g(g_), a(a_), prefs(p), rolled(r), scratch(s) { gm_base = g.P; }

    // a constant of the run: baked in when the generator was given the values, a load from HBM otherwise
    std::string cst(int idx, Sink& s) {
        static const char* bake_env = getenv("HVB_GEN_BAKE");             // experiments: 0 = never, 1 = straight-line code only, 2 = always
        const int bake = bake_env ? atoi(bake_env) : 1;                   // measured on c5: baking into the row tables thrashes the constant cache
        if ((size_t)idx * 2 + 1 < g.run_consts.size() && (bake == 2 || (bake == 1 && !s.rolled))) {
            const double re = g.run_consts[2 * (size_t)idx], im = g.run_consts[2 * (size_t)idx + 1];
            return "cmk(" + s.d(re) + ", " + (im == 0.0 ? std::string("0.0") : s.d(im)) + ")";
        }
        return "D.pv.consts[" + s.i(idx) + "]";
    }

    // parameter pi is a constant of the given kind whose value in this run has no imaginary part
    bool const_is_real(int pi, int kind) const {
        if (prefs[2 * pi] != kind) return false;
        const size_t idx = (size_t)prefs[2 * pi + 1];
        return idx * 2 + 1 < g.run_consts.size() && g.run_consts[2 * idx + 1] == 0.0 && std::isfinite(g.run_consts[2 * idx]);
    }

    std::string param(int pi, Sink& s) {
        const int kind = prefs[2 * pi], idx = prefs[2 * pi + 1];
        switch (kind) {
            case PK_CONST: return cst(idx, s);
            case PK_SAMPLE: return "cmk(srow[" + s.i(idx) + "], 0.0)";
            case PK_SAMPLE_NEG: return "cmk(-srow[" + s.i(idx) + "], 0.0)";
            case PK_SAMPLE_INV: return "cmk(1.0 / srow[" + s.i(idx) + "], 0.0)";
            case PK_TABLE: return "D.pv.tables[(long long)" + s.i(idx) + " * D.pv.table_ld + fi]";
            case PK_SPLINE: return "spline_eval(D.pv.spl_t, D.pv.spl_c, D.pv.spl_trace, D.pv.spl_fill, D.pv.spl_range, " + s.i(idx) + ", f)";
            case PK_SMD: return "eval_smd(D.pv.consts + " + s.i(idx) + ", f)";
            case PK_BETA_MUL: return "cmk(__dmul_rn(w, D.pv.consts[" + s.i(idx) + "].x) / HVB_C0, 0.0)";
            case PK_BETA: return "gen_beta(wc0, " + cst(idx, s) + ")";
        }
        return "cmk(0.0, 0.0)";
    }

    // code that evaluates the value ops first used at row r; registers their variable names
    std::string value_code(int r, Sink& s) {
        std::string t;
        const int p = r & 1;
        int ns = 0, nc = 0, nt = 0, nl = 0;
        for (int o = 0; o < g.nfast; ++o) {
            if (a.fop_first[o] != r || a.fop_last[o] < 0) continue;
            const int code = g.fop_code[o];
            const std::string name = fmt("vs%d_%d", p, ns++);
            std::string arg = (code & 4) ? "srow[" + s.i((int)g.fop_arg[o]) + "]" : s.d(g.fop_arg[o]);
            if (code & 2) arg = "__dmul_rn(w, " + arg + ")";
            if (code & 1) arg = ((code & 2) ? "-rcp_fast(" : "rcp_fast(") + arg + ")";      // <= ~1 ulp; the assembly dump keeps the IEEE division
            t += "    " + name + " = " + arg + ";\n";
            fop_name[o] = name;
        }
        for (int o = 0; o < g.nops; ++o) {
            if (a.op_first[o] != r || a.op_last[o] < 0) continue;
            const int32_t* op = &g.ops[(size_t)o * HVB_OP_WORDS];
            const int code = op[0];
            if (code == OP_TLINE) {
                if ((a.op_submask[o] & ~0xFu) == 0) {
                    // both terminals' reference is ground: only y11 = y22 and y12 = y21 are stamped (det(ABCD) = 1)
                    const std::string name = fmt("vl%d_%d", p, nl++);
                    op_name[o] = name;
                    op_chain_form[o] = true;
                    // the kernel is compiled for this run's constants (the library keys its variants on them): a line whose
                    // Z0 and propagation constant are real constants gets the lossless form, no cosh / sinh at all
                    op_imag_only[o] = const_is_real(op[2], PK_CONST) && const_is_real(op[3], PK_BETA) && const_is_real(op[4], PK_CONST);
                    if (op_imag_only[o])
                        t += "    gen_tline_lossless(" + param(op[2], s) + ".x, " + param(op[3], s) + ".x, " + param(op[4], s) + ".x, " + name + "a.y, " + name + "b.y);\n";
                    else
                        t += "    gen_tline_chain(" + param(op[2], s) + ", " + param(op[3], s) + ", " + param(op[4], s) + ".x, " + name + "a, " + name + "b);\n";
                    continue;
                }
                const std::string name = fmt("vt%d_%d", p, nt++);
                t += "    eval_tline_fast(" + param(op[2], s) + ", " + param(op[3], s) + ", " + param(op[4], s) + ".x, " + name + ");\n";
                op_name[o] = name;
                op_chain_form[o] = false;
                continue;
            }
            const std::string name = fmt("vc%d_%d", p, nc++);
            const std::string p0 = param(op[2], s);
            if (code == OP_COPY) t += "    " + name + " = " + p0 + ";\n";
            else if (code == OP_RECIP) t += "    " + name + " = crecip_np(" + p0 + ");\n";
            else if (code == OP_CAP) t += "    " + name + " = gen_cap(w, " + p0 + ");\n";
            else t += "    " + name + " = gen_ind(w, " + p0 + ");\n";
            op_name[o] = name;
        }
        max_vs[p] = std::max(max_vs[p], ns); max_vc[p] = std::max(max_vc[p], nc); max_vt[p] = std::max(max_vt[p], nt);
        max_vl[p] = std::max(max_vl[p], nl);
        return t;
    }

    ValRef value(int vid, Sink& s) {
        if (vid == HVB_GEN_VID_ONE) return ValRef{"1.0", "0"};
        if (vid < g.nvc) {
            const double re = g.const_vals[2 * (size_t)vid], im = g.const_vals[2 * (size_t)vid + 1];
            return ValRef{re == 0.0 ? "0" : s.d(re), im == 0.0 ? "0" : s.d(im)};
        }
        const Prod& p = a.prod[vid - g.nvc];
        if (p.kind == 0) {
            const std::string& nm = fop_name[p.idx];
            return (g.fop_code[p.idx] & 2) ? ValRef{"0", nm} : ValRef{nm, "0"};
        }
        const std::string& nm = op_name[p.idx];
        const bool line = g.ops[(size_t)p.idx * HVB_OP_WORDS] == OP_TLINE;
        if (line && op_chain_form[p.idx]) {
            const std::string b2 = nm + ((p.sub == 0 || p.sub == 3) ? "a" : "b");
            return ValRef{op_imag_only[p.idx] ? std::string("0") : b2 + ".x", b2 + ".y"};
        }
        const std::string base = line ? nm + fmt("[%d]", p.sub) : nm;
        return ValRef{base + ".x", base + ".y"};
    }

    // one matrix / right-hand-side cell: the entries' values summed in component order (the reference's `+=`)
    std::string cell(int r, int col, Sink& s) {
        std::string re, im;
        auto add = [](std::string& acc, const std::string& term, bool neg) {
            if (term == "0") return;
            if (acc.empty()) acc = neg ? "-" + term : term;
            else acc += (neg ? " - " : " + ") + term;
        };
        for (const Entry& en : a.rows[r]) {
            if (en.col != col) continue;
            const ValRef v = value(en.vid, s);
            add(re, v.re, en.neg);
            add(im, v.im, en.neg);
        }
        return "cmk(" + (re.empty() ? std::string("0.0") : re) + ", " + (im.empty() ? std::string("0.0") : im) + ")";
    }

    // assemble row r: into the current row (r == 0) or the incoming row (r >= 1)
    std::string assemble(int r, Sink& s) {
        std::string t = value_code(r, s);
        const int n = g.n;
        if (r == 0) {
            t += "    c0 = " + cell(0, 0, s) + ";\n";
            t += "    c1 = " + (n > 1 ? cell(0, 1, s) : std::string("cmk(0.0, 0.0)")) + ";\n";
            t += "    c2 = cmk(0.0, 0.0);\n";
            if (a.tap_of_row[0] >= 0) t += "    " + GM(0) + " = " + cell(0, n + a.tap_port[0], s) + ";\n";
        } else {
            t += "    n0 = " + cell(r, r - 1, s) + ";\n";
            t += "    n1 = " + cell(r, r, s) + ";\n";
            t += "    n2 = " + (r + 1 < n ? cell(r, r + 1, s) : std::string("cmk(0.0, 0.0)")) + ";\n";
            if (a.tap_of_row[r] >= 0) t += "    gf = " + cell(r, n + a.tap_port[a.tap_of_row[r]], s) + ";\n";
            if (g.cut_col >= 0 && r == n - 1) t += "    cutc = " + cell(r, g.cut_col, s) + ";\n";      // coupling to the other half
        }
        // values last used here die; forget their names so a stale reference cannot compile silently
        return t;
    }

    std::string slot(int i) const { return fmt("gsm[%d * GEN_THREADS + threadIdx.x]", i); }
    std::string acc_ref(int o, int j) const {                     // accumulator of (output tap o, source tap j)
        if (half) {
            // real pairs sit in their S cell; a real row's extra-tap column (alpha_o) in the row's cell of a port of the
            // other half; the extra tap's own row (beta_j, delta) in shared memory
            const int go = gport[a.tap_port[o]], gj = gport[a.tap_port[j]];
            if (go < 0) return slot(x_base + j);
            return fmt("Sp[%dLL * D.ldS]", go * Pglob + (gj < 0 ? stash : gj));
        }
        if (!scratch) return fmt("ACC_%d_%d", o, j);
        return fmt("Sp[%dLL * D.ldS]", a.tap_port[o] * g.P + a.tap_port[j]);
    }
    // per-port running sums T_o and column factors gamma_j: registers for few ports, shared memory otherwise
    std::string TT(int o) const { return scratch ? slot(tt_base + o) : fmt("TT_%d", o); }
    std::string GM(int j) const { return scratch ? slot(gm_base + j) : fmt("GM_%d", j); }

    // fold gamma_j * T_o of the epoch that just ended into the pair accumulators; f = taps that are old now
    std::string flush(int f, bool final_) {
        std::string t;
        if (f > 0) {
            t += "    {\n";
            for (int j = 0; j < f; ++j) t += fmt("      const cplx gm%d = %s;\n", j, GM(j).c_str());
        }
        for (int o = 0; o < f; ++o) {
            t += fmt("      { const cplx tt = %s;\n", TT(o).c_str());
            for (int j = 0; j < f; ++j) {
                const bool first = scratch && o == f - 1;              // first touch of (o, j): see the header comment
                const std::string prodt = fmt("cmul(gm%d, tt)", j);
                if (final_) {
                    t += fmt("        { cplx x = %s; ", prodt.c_str());
                    if (!first) t += "const cplx x0 = " + acc_ref(o, j) + "; x.x += x0.x; x.y += x0.y; ";
                    const int po = a.tap_port[o], pj = a.tap_port[j];
                    t += fmt("x.x = fma(2.0, x.x, %s); x.y = 2.0 * x.y; ", po == pj ? "-1.0" : "0.0");
                    const double sc = g.epi_scale[(size_t)po * g.P + pj];
                    if (sc != 1.0) t += fmt("x = cscale(x, %a); ", sc);
                    t += "if (!(isfinite(x.x) && isfinite(x.y))) bad |= HVB_STATUS_NONFINITE_BIT; ";
                    t += fmt("Sp[%dLL * D.ldS] = x; }\n", po * g.P + pj);
                } else if (first) {
                    t += "        " + acc_ref(o, j) + " = " + prodt + ";\n";
                } else if (scratch) {
                    t += "        { cplx x = " + acc_ref(o, j) + fmt("; cfma(x, gm%d, tt); ", j) + acc_ref(o, j) + " = x; }\n";
                } else {
                    t += fmt("        cfma(ACC_%d_%d, gm%d, tt);\n", o, j, j);
                }
                flops += 8;
            }
            if (!final_) t += "        " + TT(o) + " = cmk(0.0, 0.0);\n";
            t += "      }\n";
        }
        if (f > 0) t += "    }\n";
        return t;
    }

    // elimination step k: par = k & 1, nb = ports whose row is <= k, newly = row k is a port row, fresh = row
    // k+1 is a port row (its source column joins the old ones after this step), last = no row k+1.
    // w_o[k-1] / w_o[k-2] alternate between WA_o and WB_o, U[k-2][k] between GA and GB: no register moves.
    std::string step(int par, int nb, bool newly, bool fresh, bool last) {
        std::string t;
        const bool old = nb > 0;                                       // old source columns exist (ports at rows <= k)
        const char* Wprev = par ? "WB" : "WA";                         // w[k-1]
        const char* Wold = par ? "WA" : "WB";                          // w[k-2], overwritten by w[k]
        const char* Gold = par ? "GB" : "GA";                          // U[k-2][k], overwritten by U[k][k+2]
        if (last) {
            t += "    { double best = fabs(c0.x) + fabs(c0.y); if (!(best >= 0.0)) best = -1.0;\n";
            t += "      pmin = fmin(pmin, best); pmax = fmax(pmax, best);\n";
            t += "      const cplx p0 = c0;\n";
        } else {
            t += "    { double best = fabs(c0.x) + fabs(c0.y); if (!(best >= 0.0)) best = -1.0;\n";
            t += "      const double mn = fabs(n0.x) + fabs(n0.y); const bool sw = mn > best; if (sw) best = mn;\n";
            t += "      pmin = fmin(pmin, best); pmax = fmax(pmax, best);\n";
            t += "      const cplx p0 = sw ? n0 : c0, p1 = sw ? n1 : c1, p2 = sw ? n2 : c2;\n";
            t += "      const cplx r0 = sw ? c0 : n0; cplx r1 = sw ? c1 : n1, r2 = sw ? c2 : n2;\n";
        }
        t += "      const cplx inv = crecip_fast(p0);\n";
        flops += 8;
        if (!last) {
            t += "      const cplx m = cmul(r0, inv); cfnma(r1, m, p1); cfnma(r2, m, p2);\n";
            flops += 6 + 16;
        }
        if (old) {
            if (last) t += "      const cplx zeta = z;\n";
            else {
                t += "      const cplx zeta = sw ? cmk(0.0, 0.0) : z;\n";
                t += "      const cplx znext = sw ? z : cneg(cmul(m, z));\n";
                flops += 6;
            }
        }
        if (fresh) {
            t += "      const cplx yf = sw ? gf : cmk(0.0, 0.0);\n";
            t += "      const cplx rf = sw ? cneg(cmul(m, gf)) : gf;\n";
            flops += 6;
        }
        if (nb > (newly ? 1 : 0)) {
            t += fmt("      const cplx ninv = cneg(inv), e1 = cmul(h1, ninv), e2 = cmul(%s, ninv);\n", Gold);
            flops += 12;
        }
        for (int o = 0; o < nb; ++o) {
            if (newly && o == nb - 1) t += "      { const cplx wk = inv;\n";
            else {
                t += fmt("      { cplx wk = cmul(%s_%d, e1); cfma(wk, %s_%d, e2);\n", Wprev, o, Wold, o);
                flops += 6 + 8;
            }
            if (newly && o == nb - 1) t += "        " + TT(o) + " = cmul(wk, zeta);\n";
            else if (scratch) t += "        { cplx tt = " + TT(o) + "; cfma(tt, wk, zeta); " + TT(o) + " = tt; }\n";
            else t += "        cfma(" + TT(o) + ", wk, zeta);\n";
            flops += 8;
            if (fresh) { t += "        " + acc_ref(o, nb) + " = cmul(wk, yf);\n"; flops += 6; }
            t += fmt("        %s_%d = wk; }\n", Wold, o);
        }
        if (!last) {
            t += fmt("      %s = p2; h1 = p1;\n", Gold);
            t += "      c0 = r1; c1 = r2; c2 = cmk(0.0, 0.0);\n";
            if (old) t += "      z = znext;\n";
        }
        if (fresh) {
            // the source column of port nb becomes an old column: close the epoch
            t += flush(nb, false);
            for (int j = 0; j < nb; ++j) { t += "      " + GM(j) + " = cmul(" + GM(j) + ", z);\n"; flops += 6; }
            t += "      " + GM(nb) + " = rf; z = cmk(1.0, 0.0);\n";
        }
        t += "    }\n";
        return t;
    }
};

// grid-stride loop over the points and the per-point preamble (frequency, sample row, output cell)
static std::string point_loop_open(int P) {
    std::string src;
    src += "  const long long T = (long long)gridDim.x * blockDim.x;\n";
    src += "  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n";
    src += "  const long long total = D.nS * D.nF;\n  unsigned bad = 0;\n";
    src += "  for (long long q = t0; q < total; q += T) {\n";
    src += "    long long s = 0, fi = q;\n";
    src += "    if (D.nS != 1) {                                   // one sweep: no division at all\n";
    src += "      if (total < 0x7fffffffLL) { const unsigned uq = (unsigned)q, un = (unsigned)D.nF; s = uq / un; fi = uq - (unsigned)s * un; }\n";
    src += "      else { s = q / D.nF; fi = q - s * D.nF; }\n    }\n";
    src += "    const double f = __ldg(D.f + fi);\n";
    src += "    if (D.f32 && s == 0) D.f32[fi] = (float)f;\n";
    src += "    const double* __restrict__ srow = D.samples + s * D.nR; (void)srow;\n";
    src += "    const double w = __dmul_rn(HVB_TWOPI, f); (void)w;\n";
    src += "    const double wc0 = w / HVB_C0; (void)wc0;\n";
    src += fmt("    cplx* __restrict__ Sp = D.S + s * %dLL * D.ldS + fi;\n", P * P);
    src += "    double pmin = 1e300, pmax = 0.0;\n";
    return src;
}

static std::string helper_text(bool rolled) {
    std::string src;
    src += "__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {\n"
           "    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);\n"
           "    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);\n}\n"
           "__device__ __forceinline__ cplx gen_beta(double b0, cplx s) {          // b0 = (2 pi f) / c0, once per point\n"
           "    return cmk(__dmul_rn(b0, s.x), s.y == 0.0 ? 0.0 : __dmul_rn(b0, s.y));\n}\n"
           "__device__ __forceinline__ cplx gen_cap(double w, cplx p0) {\n"
           "    return (p0.y == 0.0) ? cmk(0.0, __dmul_rn(w, p0.x)) : cmk(-__dmul_rn(w, p0.y), __dmul_rn(w, p0.x));\n}\n"
           "__device__ __forceinline__ cplx gen_ind(double w, cplx p0) {\n"
           "    if (p0.y == 0.0) return cmk(0.0, -rcp_fast(__dmul_rn(w, p0.x)));\n"
           "    return cdiv_np(cmk(1.0, 0.0), cmk(-__dmul_rn(w, p0.y), __dmul_rn(w, p0.x)));\n}\n"
           "// line between two nodes whose reference is ground (base/tl.py:30-60): a11 = a22 = cosh(th), a12 = Z0 sinh(th),\n"
           "// and det(ABCD) = cosh^2 - sinh^2 = 1, so y11 = y22 = a11 / a12 and y12 = y21 = -1 / a12\n"
           "// lossless line with real Z0: cosh(th) = cos(y), sinh(th) = i sin(y), so y11 = -i cos(y) / (Z0 sin(y)), y12 = i / (Z0 sin(y))\n"
           "__device__ __forceinline__ void gen_tline_lossless(double Z0, double beta, double len, double& ya_im, double& yb_im) {\n"
           "    double sy, cy;\n"
           "    sincos(beta * len, &sy, &cy);\n"
           "    const double r = rcp_fast(Z0 * sy);\n"
           "    ya_im = -(cy * r);\n"
           "    yb_im = r;\n}\n";
    // a row loop carries one copy of the general form instead of one per row shape (instruction-cache footprint)
    src += std::string("__device__ ") + (rolled ? "__noinline__" : "__forceinline__") + " void gen_tline_chain(cplx Z0, cplx beta, double len, cplx& ya, cplx& yb) {\n"
           "    const double x = -beta.y * len, y = beta.x * len;\n"
           "    double sy, cy;\n"
           "    sincos(y, &sy, &cy);\n"
           "    cplx ch, sh;\n"
           "    if (x == 0.0) { ch = cmk(cy, 0.0); sh = cmk(0.0, sy); }\n"
           "    else { const double chx = cosh(x), shx = sinh(x); ch = cmk(chx * cy, shx * sy); sh = cmk(shx * cy, chx * sy); }\n"
           "    const cplx r = crecip_fast(cmul(Z0, sh));\n"
           "    ya = cmul(ch, r);\n"
           "    yb = cneg(r);\n}\n\n";
    return src;
}

// the row tables as __constant__ initialisers of the module
static std::string tables_text(const GenResult& res, int nrows) {
    const GenResult* out = &res;
    std::string src;
    src += fmt("__constant__ int4 gen_rw[%d] = {", std::max(nrows, 1));
    for (size_t k = 0; k + 3 < out->rows.size(); k += 4) src += fmt("%s{%d,%d,%d,%d}", k ? "," : "", out->rows[k], out->rows[k + 1], out->rows[k + 2], out->rows[k + 3]);
    src += "};\n";
    src += fmt("__constant__ int gen_it[%d] = {", (int)std::max<size_t>(out->itab.size(), 1));
    for (size_t k = 0; k < out->itab.size(); ++k) src += fmt("%s%d", k ? "," : "", out->itab[k]);
    if (out->itab.empty()) src += "0";
    src += "};\n";
    src += fmt("__constant__ double gen_dt[%d] = {", (int)std::max<size_t>(out->dtab.size(), 1));
    for (size_t k = 0; k < out->dtab.size(); ++k) {
        const double v = out->dtab[k];
        src += k ? "," : "";
        src += (std::isfinite(v) ? fmt("%a", v) : std::string(std::isnan(v) ? "__longlong_as_double(0x7ff8000000000000LL)" : (v > 0 ? "__longlong_as_double(0x7ff0000000000000LL)" : "__longlong_as_double(0xfff0000000000000LL)")));
    }
    if (out->dtab.empty()) src += "0.0";
    src += "};\n\n";
    return src;
}

// ---- two-sided form (see the header comment) ------------------------------------------------------------------
struct HalfPlan {
    GenPlan g;
    std::vector<int> gport;            // local port -> port of the whole plan; the last one (-1) is the extra tap
    int nreal = 0;
};

// rows: the half's rows in walking order (indices of the whole plan); cut: the row on the other side of the cut
static bool make_half(const GenPlan& g, const std::vector<int>& rows, int cut, HalfPlan* h, std::string* why) {
    const int n = g.n, nh = (int)rows.size();
    auto no = [&](const char* w) { if (why) *why = w; return false; };
    std::vector<int> loc(n, -1);
    for (int i = 0; i < nh; ++i) loc[rows[i]] = i;
    GenPlan& o = h->g;
    o = GenPlan();
    o.n = nh; o.nvc = g.nvc; o.ndyn = g.ndyn; o.nops = g.nops; o.nfast = g.nfast; o.ncoops = 0; o.epi_simple = 1; o.nprefs = g.nprefs;
    o.const_vals = g.const_vals; o.ops = g.ops; o.fop_code = g.fop_code; o.fop_out = g.fop_out; o.fop_arg = g.fop_arg;
    o.run_consts = g.run_consts; o.spl_trace = g.spl_trace;
    std::vector<int> lport(g.P, -1);
    h->gport.clear();
    for (int q = 0; q < g.P; ++q)
        if (loc[g.port_rows[3 * q]] >= 0) { lport[q] = (int)h->gport.size(); h->gport.push_back(q); }
    h->nreal = (int)h->gport.size();
    h->gport.push_back(-1);
    if (h->nreal < 1) return no("a half chain without a port");
    o.P = h->nreal + 1; o.ncols = nh + o.P; o.cut_col = nh + o.P;
    if (o.cut_col > 0x7FF) return no("half chain too long for the entry format");
    o.port_rows.assign(3 * (size_t)o.P, -1);
    o.epi_port_of_row.assign(nh, -1);
    for (int lq = 0; lq < h->nreal; ++lq) {
        const int r = loc[g.port_rows[3 * h->gport[lq]]];
        o.port_rows[3 * lq] = r; o.epi_port_of_row[r] = lq;
    }
    if (o.epi_port_of_row[nh - 1] >= 0) return no("cut next to a port row");
    o.port_rows[3 * h->nreal] = nh - 1; o.epi_port_of_row[nh - 1] = h->nreal;
    o.epi_scale.assign((size_t)o.P * o.P, 1.0);
    o.row_ptr.assign(1, 0);
    for (int i = 0; i < nh; ++i) {
        const int r = rows[i];
        for (int e = g.row_ptr[r]; e < g.row_ptr[r + 1]; ++e) {
            const uint32_t w = (uint32_t)g.row_ent[e];
            const int col = (int)((w >> 20) & 0x7FFu);
            int c2;
            if (col >= n) {
                const int q = col - n;
                if (q >= g.P || lport[q] < 0) return no("source entry of a port of the other half");
                c2 = nh + lport[q];
            } else if (loc[col] >= 0) c2 = loc[col];
            else if (col == cut && i == nh - 1) c2 = o.cut_col;
            else return no("row coupled across the cut");
            o.row_ent.push_back((int32_t)(((uint32_t)c2 << 20) | (w & 0xFFFFFu)));
        }
        if (i == nh - 1) o.row_ent.push_back((int32_t)(((uint32_t)(nh + h->nreal) << 20) | ((uint32_t)HVB_GEN_VID_ONE << 1)));
        o.row_ptr.push_back((int32_t)o.row_ent.size());
    }
    return true;
}

// the cut m | m+1 that minimises the border work (tap-rows); -1 when the chain offers none.  *ratio = work / one-sided work
static int choose_cut(const GenPlan& g, const Analysis& a, double* ratio) {
    const int n = g.n;
    long long one = 0;
    for (int r : a.tap_row) one += n - r;
    int best = -1;
    long long bestc = 0;
    for (int m = 1; m + 3 <= n; ++m) {                                     // front rows 0..m, back rows m+1..n-1: two rows each at least
        if (a.tap_of_row[m] >= 0 || a.tap_of_row[m + 1] >= 0) continue;
        long long c = 2;
        int nf = 0, nb = 0;
        for (int r : a.tap_row) { if (r < m) { c += m - r + 1; ++nf; } else { c += r - m; ++nb; } }
        if (!nf || !nb) continue;
        if (best < 0 || c < bestc || (c == bestc && std::abs(2 * m + 1 - n) < std::abs(2 * best + 1 - n))) { best = m; bestc = c; }
    }
    if (ratio) *ratio = (best >= 0 && one > 0) ? (double)bestc / (double)one : 1.0;
    return best;
}

static int gen_two_sided(const GenPlan& g, const int32_t* prefs, const char* prelude, int m, GenResult* out, std::string* err) {
    const int n = g.n, P = g.P;
    std::vector<int> fr, bk;
    for (int r = 0; r <= m; ++r) fr.push_back(r);
    for (int r = n - 1; r > m; --r) bk.push_back(r);
    HalfPlan hf, hb;
    if (!make_half(g, fr, m + 1, &hf, err) || !make_half(g, bk, m, &hb, err)) return -1;
    const Analysis af = analyse(hf.g), ab = analyse(hb.g);
    if (!af.ok) { if (err) *err = "front half: " + af.why; return -1; }
    if (!ab.ok) { if (err) *err = "back half: " + ab.why; return -1; }
    const int nF = hf.g.n, nB = hb.g.n, PF = hf.g.P, PB = hb.g.P, PM = std::max(PF, PB);
    Gen GF(hf.g, af, prefs, true, true), GK(hb.g, ab, prefs, true, true);
    for (Gen* G : {&GF, &GK}) { G->half = true; G->Pglob = P; G->tt_base = 0; G->gm_base = PM; }
    GF.gport = hf.gport; GK.gport = hb.gport;
    GF.stash = hb.gport[0]; GK.stash = hf.gport[0];
    GF.x_base = 2 * PM; GK.x_base = 3 * PM + 1;
    const int vslot = GF.x_base + PF;                                      // v = A[m][m+1] waits here while the back half is walked
    const int nslots = 4 * PM + 1;
    int threads = (nslots * 64 * 16 > 48 * 1024) ? 32 : 64;               // measured on c5: 64 x 6 CTAs 2.87 ms, 128 x 3 3.03 ms
    if (const char* e = getenv("HVB_GEN_THREADS")) { if (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 96 || atoi(e) == 128) threads = atoi(e); }   // experiments
    if (nslots * threads * 16 > 48 * 1024) { if (err) *err = "too many ports for the shared-memory sums"; return -1; }

    out->rolled = true; out->scratch_acc = true; out->two_sided = true; out->cut_row = m;
    out->rows.clear(); out->itab.clear(); out->dtab.clear();
    auto nbs = [](const Analysis& a, int nh) { std::vector<int> nb(nh, 0); int c = 0; for (int k = 0; k < nh; ++k) { if (a.tap_of_row[k] >= 0) ++c; nb[k] = c; } return nb; };
    const std::vector<int> nbF = nbs(af, nF), nbB = nbs(ab, nB);

    Sink s0{false, &out->itab, &out->dtab};
    const std::string head = GF.assemble(0, s0);
    std::vector<std::string> asm_cases, step_cases;
    std::map<std::string, int> asm_id, step_id;
    auto add_row = [&](const std::string& at, const std::string& st, int ib, int db) {
        auto ia = asm_id.find(at);
        int ca, cs;
        if (ia == asm_id.end()) { ca = (int)asm_cases.size(); asm_id[at] = ca; asm_cases.push_back(at); } else ca = ia->second;
        auto is = step_id.find(st);
        if (is == step_id.end()) { cs = (int)step_cases.size(); step_id[st] = cs; step_cases.push_back(st); } else cs = is->second;
        out->rows.push_back(ca); out->rows.push_back(cs); out->rows.push_back(ib); out->rows.push_back(db);
    };
    auto walk = [&](Gen& G, const Analysis& a, const std::vector<int>& nb, int nh) {
        for (int k = 0; k + 1 < nh; ++k) {
            Sink s{true, &out->itab, &out->dtab};
            const int ib = (int)out->itab.size(), db = (int)out->dtab.size();
            const std::string at = G.assemble(k + 1, s);
            const std::string st = G.step(k & 1, nb[k], a.tap_of_row[k] >= 0, a.tap_of_row[k + 1] >= 0, false);
            add_row(at, st, ib, db);
        }
    };
    walk(GF, af, nbF, nF);
    {
        // turn-around: the front half's last column, its sums, and the back half's first row (row n-1)
        Sink s{true, &out->itab, &out->dtab};
        const int ib = (int)out->itab.size(), db = (int)out->dtab.size();
        std::string st = GF.step((nF - 1) & 1, nbF[nF - 1], af.tap_of_row[nF - 1] >= 0, false, true);
        st += GF.flush(PF, false);
        st += "    " + GF.slot(vslot) + " = cutc;\n";
        st += "    z = cmk(1.0, 0.0); h1 = cmk(0.0, 0.0); GA = h1; GB = h1;\n";
        for (int o = 0; o < PM; ++o) st += fmt("    WA_%d = h1; WB_%d = h1;\n", o, o);
        st += GK.assemble(0, s);
        add_row("", st, ib, db);
    }
    walk(GK, ab, nbB, nB);
    const int nrows = (int)out->rows.size() / 4;
    const size_t table_bytes = out->rows.size() * 4 + out->itab.size() * 4 + out->dtab.size() * 8;
    if (table_bytes > 56 * 1024) { if (err) *err = "row tables exceed the constant bank"; return -1; }
    out->tables_in_source = true;

    std::string body;
    body += "    int4 rwn = gen_rw[0];\n";
    body += fmt("    for (int k = 0; k < %d; ++k) {\n", nrows);
    body += fmt("      const int4 rw = rwn; rwn = gen_rw[k + 1 < %d ? k + 1 : k];          // next row's entry is fetched one row ahead\n", nrows);
    body += "      const int* __restrict__ it = gen_it + rw.z; const double* __restrict__ dt = gen_dt + rw.w; (void)it; (void)dt;\n";
    body += "      switch (rw.x) {\n";
    for (size_t c = 0; c < asm_cases.size(); ++c) body += fmt("      case %d: {\n", (int)c) + asm_cases[c] + "      } break;\n";
    body += "      }\n      switch (rw.y) {\n";
    for (size_t c = 0; c < step_cases.size(); ++c) body += fmt("      case %d:\n", (int)c) + step_cases[c] + "      break;\n";
    body += "      }\n    }\n";
    out->n_cases = (int)(asm_cases.size() + step_cases.size());
    body += "    // ---- the back half's last column (row m+1) and its sums\n";
    body += GK.step((nB - 1) & 1, nbB[nB - 1], ab.tap_of_row[nB - 1] >= 0, false, true);
    body += GK.flush(PB, false);

    // ---- interface system, port voltages -> S (solvers/spfn.py:159-176, closed form for ground-referenced ports)
    long long cflops = 0;
    auto tap_of_port = [](const Analysis& a, const HalfPlan& h, int gq) {       // tap index of whole-plan port gq in this half, or -1
        for (size_t t = 0; t < a.tap_port.size(); ++t) if (h.gport[a.tap_port[t]] == gq) return (int)t;
        return -1;
    };
    std::string c;
    c += "    // ---- the cut: 2 x 2 interface system, then the port voltages of both halves\n";
    c += "    double canc;\n    {\n";
    c += "      const cplx dF = " + GF.slot(GF.x_base + PF - 1) + ", dB = " + GK.slot(GK.x_base + PB - 1) + ", cv = " + GF.slot(vslot) + ", cu = cutc;\n";
    c += "      const cplx kap = cmul(cv, dF), udB = cmul(cu, dB);\n";
    c += "      cplx Dl = cmk(1.0, 0.0); cfnma(Dl, kap, udB);\n";
    c += "      canc = (1.0 + (fabs(kap.x) + fabs(kap.y)) * (fabs(udB.x) + fabs(udB.y))) / (fabs(Dl.x) + fabs(Dl.y));\n";
    c += "      const cplx iD = crecip_fast(Dl);\n";
    cflops += 6 + 6 + 8 + 8 + 8;
    for (int gq = 0; gq < P; ++gq) {
        const int tf = tap_of_port(af, hf, gq), tb = tap_of_port(ab, hb, gq);
        if (tf >= 0) {
            c += fmt("      const cplx xi_%d = cmul(%s, iD); const cplx eta_%d = cneg(cmul(udB, xi_%d));\n", gq, GF.slot(GF.x_base + tf).c_str(), gq, gq);
            cflops += 12;
        } else {
            c += fmt("      const cplx bb_%d = %s; const cplx xi_%d = cneg(cmul(cmul(kap, bb_%d), iD)); cplx eta_%d = bb_%d; cfnma(eta_%d, udB, xi_%d);\n",
                     gq, GK.slot(GK.x_base + tb).c_str(), gq, gq, gq, gq, gq, gq);
            cflops += 20;
        }
    }
    for (int side = 0; side < 2; ++side) {
        const Analysis& a = side ? ab : af;
        const HalfPlan& h = side ? hb : hf;
        const Gen& G = side ? GK : GF;
        for (int t = 0; t < h.nreal; ++t) {                                    // real taps come first (the extra tap is on the last row)
            const int go = h.gport[a.tap_port[t]];
            c += fmt("      { const cplx al = cmul(Sp[%dLL * D.ldS], %s);\n", go * P + G.stash, side ? "cu" : "cv");
            cflops += 6;
            for (int gq = 0; gq < P; ++gq) {
                const bool own = tap_of_port(a, h, gq) >= 0;
                const double sc = g.epi_scale[(size_t)go * P + gq];
                c += fmt("        { const cplx t2 = cmul(al, %s_%d); ", side ? "xi" : "eta", gq);
                if (own) {
                    c += fmt("cplx x = Sp[%dLL * D.ldS]; ", go * P + gq);
                    c += fmt("canc = fmax(canc, (fabs(x.x) + fabs(x.y) + fabs(t2.x) + fabs(t2.y)) * %a); x.x -= t2.x; x.y -= t2.y; ", 2.0 * std::fabs(sc));
                } else c += "cplx x = cneg(t2); ";
                c += fmt("x.x = fma(2.0, x.x, %s); x.y = 2.0 * x.y; ", go == gq ? "-1.0" : "0.0");
                if (sc != 1.0) c += fmt("x = cscale(x, %a); ", sc);
                c += "if (!(isfinite(x.x) && isfinite(x.y))) canc = 1e300; ";
                c += fmt("Sp[%dLL * D.ldS] = x; }\n", go * P + gq);
                cflops += 6 + 2 + 4 + (own ? 6 : 0);
            }
            c += "      }\n";
        }
    }
    c += "    }\n";
    body += c;

    std::string decl;
    for (int p = 0; p < 2; ++p) {
        for (int i = 0; i < std::max(GF.max_vs[p], GK.max_vs[p]); ++i) decl += fmt("    double vs%d_%d;\n", p, i);
        for (int i = 0; i < std::max(GF.max_vc[p], GK.max_vc[p]); ++i) decl += fmt("    cplx vc%d_%d;\n", p, i);
        for (int i = 0; i < std::max(GF.max_vt[p], GK.max_vt[p]); ++i) decl += fmt("    cplx vt%d_%d[9];\n", p, i);
        for (int i = 0; i < std::max(GF.max_vl[p], GK.max_vl[p]); ++i) decl += fmt("    cplx vl%d_%da, vl%d_%db;\n", p, i, p, i);
    }
    decl += "    cplx c0, c1, c2, n0 = cmk(0.0, 0.0), n1 = n0, n2 = n0, gf = n0, z = cmk(1.0, 0.0), h1 = n0, GA = n0, GB = n0, cutc = n0;\n";
    for (int o = 0; o < PM; ++o) decl += fmt("    cplx WA_%d = cmk(0.0, 0.0), WB_%d = WA_%d;\n", o, o, o);

    out->threads = threads;
    std::string src;
    src += "// generated by heavi_b200 (csrc/hvb_gen.cpp): kernel D for one netlist, two-sided form\n";
    src += fmt("// front half: rows 0..%d (%d ports), back half: rows %d..%d walked backwards (%d ports)\n", m, hf.nreal, n - 1, m + 1, hb.nreal);
    src += prelude;
    src += fmt("\n#define GEN_THREADS %d\n", threads);
    src += helper_text(true);
    src += tables_text(*out, nrows);
    out->kernel_name = "hvb_gen_kernel";
    // Register budget (g.occ_class, chosen in ensure_gen_kernel).  0: the compiler's own allocation -- c5: 168 registers,
    // 384 threads per SM, 2.87 ms per 1 M points; the same 168 registers forced through a minimum-CTA bound cost 5 %
    // (3.02 ms).  2: capped at 168 registers (for netlists whose natural allocation is larger: fewer than 384 resident
    // threads lose more than the spills cost).  1: 128 registers, 512 threads per SM (3.45 ms per 1 M points, spills) --
    // better when the launch is only a few waves deep.
    const int per_sm = 384 / threads;                                      // CTAs of the 168-register budget
    int minb = g.occ_class == 1 ? per_sm + per_sm / 3 : (g.occ_class == 2 ? per_sm : 0);
    if (const char* e = getenv("HVB_GEN_MINB")) minb = atoi(e);
    src += fmt("extern \"C\" __global__ void __launch_bounds__(GEN_THREADS%s) hvb_gen_kernel(const HvbDev D) {\n", minb > 0 ? fmt(", %d", minb).c_str() : "");
    src += fmt("  __shared__ cplx gsm[%d * GEN_THREADS];\n", nslots);
    src += point_loop_open(P);
    src += decl;
    src += "    // ---- row 0\n" + head + body;
    src += "    if (!(pmin >= 2.2250738585072014e-308) || !(pmax <= 1.7976931348623157e308) || !(canc <= 1e3)) bad |= HVB_STATUS_RERUN_BIT;\n";
    src += "  }\n  if (bad && D.status) atomicOr(D.status, (int)bad);\n}\n";
    out->src.swap(src);
    out->flops_per_solve = GF.flops + GK.flops + cflops;
    return 0;
}
}  // namespace

bool hvb_gen_supported(const GenPlan& g, std::string* why) {
    const Analysis a = analyse(g);
    if (why) *why = a.why;
    return a.ok;
}

int hvb_gen_source(const GenPlan& g, const int32_t* prefs, int mode, const char* prelude, GenResult* out, std::string* err) {
    const Analysis a = analyse(g);
    if (!a.ok) { if (err) *err = a.why; return -1; }
    const int n = g.n, P = g.P;
    const bool rolled = mode == 2 || mode == 3 || (mode == 0 && n > HVB_GEN_UNROLL_MAX);
    if (rolled && mode != 2 && g.allow_two_sided) {
        // long chains: walk from both ends when a cut saves a quarter of the border work (mode 3: whenever there is a cut)
        static const char* ts_env = getenv("HVB_GEN_TWO_SIDED");
        double ratio = 1.0;
        const int m = (mode == 3 || !(ts_env && atoi(ts_env) == 0)) ? choose_cut(g, a, &ratio) : -1;
        if (m >= 0 && (mode == 3 || ratio <= 0.75)) {
            std::string why;
            if (gen_two_sided(g, prefs, prelude, m, out, &why) == 0) return 0;
        }
    }
    const bool scratch = P > HVB_GEN_REG_ACC_PORTS;
    out->rolled = rolled;
    out->scratch_acc = scratch;
    out->two_sided = false; out->cut_row = -1;
    out->rows.clear(); out->itab.clear(); out->dtab.clear();
    Gen G(g, a, prefs, rolled, scratch);

    // nb(k): ports at rows <= k
    std::vector<int> nb(n, 0);
    { int c = 0; for (int k = 0; k < n; ++k) { if (a.tap_of_row[k] >= 0) ++c; nb[k] = c; } }

    std::string body;                   // code of one point, after row 0 is assembled
    Sink s0{false, &out->itab, &out->dtab};
    std::string head = G.assemble(0, s0);            // row 0 is always literal code
    std::vector<std::string> asm_cases, step_cases;
    std::map<std::string, int> asm_id, step_id;
    bool tables_in_source = false;
    if (!rolled) {
        for (int k = 0; k + 1 < n; ++k) {
            Sink s{false, &out->itab, &out->dtab};
            body += fmt("    // ---- row %d enters, column %d is eliminated\n", k + 1, k);
            body += G.assemble(k + 1, s);
            body += G.step(k & 1, nb[k], a.tap_of_row[k] >= 0, a.tap_of_row[k + 1] >= 0, false);
        }
    } else {
        for (int k = 0; k + 1 < n; ++k) {
            Sink s{true, &out->itab, &out->dtab};
            const int ib = (int)out->itab.size(), db = (int)out->dtab.size();
            const std::string at = G.assemble(k + 1, s);
            const std::string st = G.step(k & 1, nb[k], a.tap_of_row[k] >= 0, a.tap_of_row[k + 1] >= 0, false);
            auto ia = asm_id.find(at);
            int ca, cs;
            if (ia == asm_id.end()) { ca = (int)asm_cases.size(); asm_id[at] = ca; asm_cases.push_back(at); } else ca = ia->second;
            auto is = step_id.find(st);
            if (is == step_id.end()) { cs = (int)step_cases.size(); step_id[st] = cs; step_cases.push_back(st); } else cs = is->second;
            out->rows.push_back(ca); out->rows.push_back(cs); out->rows.push_back(ib); out->rows.push_back(db);
        }
        // the row tables are constants of the plan: when they fit, they go into the module's constant bank
        // (uniform, cached reads) as static initialisers; otherwise the library uploads them (D.gen_*)
        const size_t table_bytes = out->rows.size() * 4 + out->itab.size() * 4 + out->dtab.size() * 8;
        static const char* ct_env = getenv("HVB_GEN_CONST_TABLES");        // experiments: 0 = row tables in global memory
        tables_in_source = table_bytes <= 56 * 1024 && !(ct_env && atoi(ct_env) == 0);
        static const char* pf_env = getenv("HVB_GEN_PREFETCH");
        const bool prefetch = pf_env ? atoi(pf_env) != 0 : true;
        if (tables_in_source && prefetch) body += "    int4 rwn = gen_rw[0];\n";
        body += "    for (int k = 0; k < " + std::to_string(n - 1) + "; ++k) {\n";
        if (tables_in_source && !prefetch) body += "      const int4 rw = gen_rw[k];\n";
        if (tables_in_source && prefetch)
            body += fmt("      const int4 rw = rwn; rwn = gen_rw[k + 1 < %d ? k + 1 : k];          // next row's entry is fetched one row ahead\n", n - 1);
        if (tables_in_source) {
            body += "      const int* __restrict__ it = gen_it + rw.z; const double* __restrict__ dt = gen_dt + rw.w; (void)it; (void)dt;\n";
        } else {
            body += "      const int4 rw = __ldg(reinterpret_cast<const int4*>(D.gen_rows) + k);\n";
            body += "      const int* __restrict__ it = D.gen_itab + rw.z; const double* __restrict__ dt = D.gen_dtab + rw.w; (void)it; (void)dt;\n";
        }
        body += "      switch (rw.x) {\n";
        for (size_t c = 0; c < asm_cases.size(); ++c) body += fmt("      case %d: {\n", (int)c) + asm_cases[c] + "      } break;\n";
        body += "      }\n      switch (rw.y) {\n";
        for (size_t c = 0; c < step_cases.size(); ++c) body += fmt("      case %d:\n", (int)c) + step_cases[c] + "      break;\n";
        body += "      }\n    }\n";
        out->n_cases = (int)(asm_cases.size() + step_cases.size());
    }
    body += "    // ---- last column\n";
    body += G.step((n - 1) & 1, nb[n - 1], a.tap_of_row[n - 1] >= 0, false, true);
    body += "    // ---- fold the last epoch, voltages -> S (solvers/spfn.py:159-176, closed form for ground-referenced ports)\n";
    body += G.flush(P, true);

    // threads per CTA: the shared-memory sums (2 P complex words per thread) must fit 48 KB of static shared memory
    const int threads = (scratch && 2 * P * 128 * 16 > 48 * 1024) ? 64 : 128;
    out->threads = threads;
    out->tables_in_source = tables_in_source;

    // ---- declarations ------------------------------------------------------------------------
    std::string decl;
    for (int p = 0; p < 2; ++p) {
        for (int i = 0; i < G.max_vs[p]; ++i) decl += fmt("    double vs%d_%d;\n", p, i);
        for (int i = 0; i < G.max_vc[p]; ++i) decl += fmt("    cplx vc%d_%d;\n", p, i);
        for (int i = 0; i < G.max_vt[p]; ++i) decl += fmt("    cplx vt%d_%d[9];\n", p, i);
        for (int i = 0; i < G.max_vl[p]; ++i) decl += fmt("    cplx vl%d_%da, vl%d_%db;\n", p, i, p, i);
    }
    decl += "    cplx c0, c1, c2, n0 = cmk(0.0, 0.0), n1 = n0, n2 = n0, gf = n0, z = cmk(1.0, 0.0), h1 = n0, GA = n0, GB = n0;\n";
    for (int o = 0; o < P; ++o) decl += fmt("    cplx WA_%d = cmk(0.0, 0.0), WB_%d = WA_%d;\n", o, o, o);
    if (!scratch) {
        for (int o = 0; o < P; ++o) decl += fmt("    cplx TT_%d = cmk(0.0, 0.0), GM_%d = TT_%d;\n", o, o, o);
        for (int o = 0; o < P; ++o)
            for (int j = 0; j < P; ++j) decl += fmt("    cplx ACC_%d_%d = cmk(0.0, 0.0);\n", o, j);
    } else {
        for (int o = 0; o < 2 * P; ++o) decl += fmt("    gsm[%d * GEN_THREADS + threadIdx.x] = cmk(0.0, 0.0);\n", o);
    }

    std::string src;
    src += "// generated by heavi_b200 (csrc/hvb_gen.cpp): kernel D for one netlist\n";
    src += prelude;
    src += fmt("\n#define GEN_THREADS %d\n", threads);
    src += helper_text(rolled);
    if (tables_in_source) src += tables_text(*out, n - 1);
    out->kernel_name = "hvb_gen_kernel";
    // registers: few ports -> everything in registers anyway; many ports -> keep four CTAs of 128 resident
    {
        int minb = scratch ? (threads == 128 ? 3 : 6) : 0;                      // measured on c5: 168 registers beat 128 (spills) and 254 (8 warps/SM)
        if (const char* e = getenv("HVB_GEN_MINB")) minb = atoi(e);          // experiments: resident CTAs the compiler must allow
        src += fmt("extern \"C\" __global__ void __launch_bounds__(GEN_THREADS%s) hvb_gen_kernel(const HvbDev D) {\n",
                   minb > 0 ? fmt(", %d", minb).c_str() : "");
    }
    if (scratch) src += fmt("  __shared__ cplx gsm[%d * GEN_THREADS];\n", 2 * P);
    src += point_loop_open(P);
    src += decl;
    src += "    // ---- row 0\n" + head + body;
    src += "    if (!(pmin >= 2.2250738585072014e-308)) bad |= HVB_STATUS_ZERO_PIVOT_BIT;\n";
    src += "    if (!(pmax <= 1.7976931348623157e308)) bad |= HVB_STATUS_NONFINITE_BIT;\n";
    src += "  }\n  if (bad && D.status) atomicOr(D.status, (int)bad);\n}\n";
    out->src.swap(src);
    out->flops_per_solve = G.flops;
    return 0;
}
