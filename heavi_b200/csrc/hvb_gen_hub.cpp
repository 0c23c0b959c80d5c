// hvb_gen_hub.cpp -- kernel E source generator: "hub" netlists, one THREAD per (sample, frequency) point.
//
// Structure served (BASELINE configs[2], an N-port Touchstone block embedded in per-port matching networks):
// ONE N-port S block (base/nport.py:5-45) whose nodes form a dense core, and chains -- path graphs of
// two-terminal parts and lines -- each hanging off one core node, ports anywhere on the chains or on core
// nodes.  The dense kernels treat such a system as a full (N+M-1)^2 matrix; here the generated code
//   1. evaluates the S_ij(f) traces (device B-splines: knot interval + basis once, a 4-term dot per trace),
//      and turns S into Y = (1/Z0)(I-S)(I+S)^-1 = (1/Z0)(2 (I+S)^-1 - I) by an IN-PLACE Gauss-Jordan inversion with
//      partial pivoting of I+S in a per-thread Pn x Pn shared-memory tile;
//   2. eliminates every chain front to back in registers (the tridiagonal recurrence of kernel D with partial
//      pivoting inside the chain), folding it into its core node's diagonal; a port's right-hand side reaches the
//      core at ONE node (its chain's, or its own), so the core's right-hand sides are P scalars in registers;
//   3. inverts the remaining Pn x Pn core matrix in place in the same tile: core voltage i of excitation q is
//      Z[i][node(q)] * b_q;
//   4. substitutes back along the chains that carry ports and forms S (solvers/spfn.py:159-176).
// Work per point: two Pn^3-sized inversions + O(chain rows), instead of (chain rows + Pn)^3.  The tile holds the
// matrix only -- 1 KB per thread for an 8-port block, so six 32-thread CTAs fit an SM (the first version kept the
// right-hand sides beside the matrix: 2 KB per thread, three CTAs).
//
// Pivoting: partial pivoting inside each chain and inside both dense solves.  The last chain row is NOT
// exchanged with its core row (that would need a spare dense row per chain); instead the step checks the
// growth |core coupling| / |pivot| and sets HVB_STATUS_ZERO_PIVOT beyond 1e4 -- the host then re-runs the sweep
// on the dense register kernel, so a flagged point never reaches the caller.
#include "hvb_gen.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>

#include "plan_format.h"

#define HUB_MAX_CORE 8
#define HUB_MAX_CHAIN_ROWS 48

namespace {

struct Entry { int col, vid; bool neg; };
struct Prod { int kind = -1, idx = 0, sub = 0; };      // 0 fast op, 1 simple op, 2 block value (sub = i*(Pn+1)+j)

std::string fmt(const char* f, ...) __attribute__((format(printf, 1, 2)));
std::string fmt(const char* f, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return buf;
}

std::string dlit(double v) {
    if (std::isnan(v)) return "__longlong_as_double(0x7ff8000000000000LL)";
    if (std::isinf(v)) return v > 0 ? "__longlong_as_double(0x7ff0000000000000LL)" : "__longlong_as_double(0xfff0000000000000LL)";
    return fmt("(%a)", v);
}

struct Chain {
    std::vector<int> rows;      // far end -> the row next to the core
    int core = -1;              // block port index of the core node it hangs off
    std::vector<int> ports;     // port indices whose signal row is on this chain
};

struct Hub {
    bool ok = false;
    std::string why;
    int Pn = 0, coop_first = 0, coop_z0 = 0, block_base = 0;
    std::vector<std::vector<Entry>> rows;
    std::vector<int> core_row, core_of_row, chain_of_row, pos_in_chain;
    std::vector<Chain> chains;
    std::vector<Prod> prod;
    bool shared_knots = false;
};

Hub analyse_hub(const GenPlan& g, const int32_t* prefs) {
    Hub h;
    const int n = g.n, P = g.P;
    auto no = [&](const std::string& w) { h.why = w; return h; };
    if (!g.epi_simple) return no("S epilogue is not the closed form");
    if (g.ncoops != 1) return no("not exactly one N-port block");
    if (g.ncols - n != P || P < 1 || P > HUB_MAX_CORE) return no("plan is padded, has no ports or more than 8 ports");
    const int32_t* co = g.coops.data();
    if (co[0] != COOP_NPORT_S) return no("N-port block is not an S block");
    h.Pn = co[2]; h.coop_first = co[3]; h.coop_z0 = co[4]; h.block_base = co[1] - g.nvc;
    const int Pn = h.Pn, ldo = Pn + 1;
    if (Pn < 2 || Pn > HUB_MAX_CORE) return no("block has more than 8 ports");
    h.rows.resize(n);
    for (int r = 0; r < n; ++r)
        for (int e = g.row_ptr[r]; e < g.row_ptr[r + 1]; ++e) {
            const uint32_t w = (uint32_t)g.row_ent[e];
            h.rows[r].push_back(Entry{(int)((w >> 20) & 0x7FFu), (int)((w >> 1) & 0x7FFFFu), (w & 1u) != 0});
        }
    // producers
    h.prod.assign(std::max(g.ndyn, 1), Prod{});
    for (int o = 0; o < g.nfast; ++o) h.prod[g.fop_out[o]] = Prod{0, o, 0};
    for (int o = 0; o < g.nops; ++o) {
        const int32_t* op = &g.ops[(size_t)o * HVB_OP_WORDS];
        const int out = op[1] - g.nvc, cnt = op[0] == OP_TLINE ? 9 : 1;
        for (int e = 0; e < cnt; ++e) h.prod[out + e] = Prod{1, o, e};
    }
    for (int e = 0; e < ldo * ldo; ++e) {
        if (h.block_base + e < 0 || h.block_base + e >= g.ndyn) return no("bad block output range");
        h.prod[h.block_base + e] = Prod{2, 0, e};
    }
    // core rows: where the block's values land; its reference node must be ground (row / column Pn unused)
    h.core_row.assign(Pn, -1);
    h.core_of_row.assign(n, -1);
    for (int r = 0; r < n; ++r)
        for (const Entry& en : h.rows[r]) {
            if (en.vid < g.nvc) continue;
            const Prod& p = h.prod[en.vid - g.nvc];
            if (p.kind != 2) continue;
            const int i = p.sub / ldo, j = p.sub % ldo;
            if (i >= Pn || j >= Pn) return no("block's reference node is not ground");
            if (en.neg || en.col >= n) return no("unexpected block entry");
            if ((h.core_row[i] >= 0 && h.core_row[i] != r) || (h.core_row[j] >= 0 && h.core_row[j] != en.col)) return no("block node appears on several rows");
            h.core_row[i] = r; h.core_row[j] = en.col;
        }
    for (int i = 0; i < Pn; ++i) {
        if (h.core_row[i] < 0) return no("block port without a row");
        if (h.core_of_row[h.core_row[i]] >= 0) return no("two block ports on one node");
        h.core_of_row[h.core_row[i]] = i;
    }
    // every (i, j) of the block must be stamped exactly once on (core_row[i], core_row[j])
    {
        std::vector<int> seen((size_t)Pn * Pn, 0);
        for (int r = 0; r < n; ++r)
            for (const Entry& en : h.rows[r]) {
                if (en.vid < g.nvc) continue;
                const Prod& p = h.prod[en.vid - g.nvc];
                if (p.kind == 2) seen[(size_t)(p.sub / ldo) * Pn + p.sub % ldo]++;
            }
        for (int v : seen) if (v != 1) return no("block values are not stamped exactly once");
    }
    // chains: connected components of the non-core rows
    std::vector<std::set<int>> adj(n);
    for (int r = 0; r < n; ++r)
        for (const Entry& en : h.rows[r])
            if (en.col < n && en.col != r) { adj[r].insert(en.col); adj[en.col].insert(r); }
    h.chain_of_row.assign(n, -1);
    h.pos_in_chain.assign(n, -1);
    for (int r0 = 0; r0 < n; ++r0) {
        if (h.core_of_row[r0] >= 0 || h.chain_of_row[r0] >= 0) continue;
        std::vector<int> comp, stack{r0};
        const int id = (int)h.chains.size();
        h.chain_of_row[r0] = id;
        while (!stack.empty()) {
            const int r = stack.back(); stack.pop_back();
            comp.push_back(r);
            for (int c : adj[r]) if (h.core_of_row[c] < 0 && h.chain_of_row[c] < 0) { h.chain_of_row[c] = id; stack.push_back(c); }
        }
        // the row next to the core, then walk the path away from it
        int near = -1, core = -1;
        for (int r : comp)
            for (int c : adj[r])
                if (h.core_of_row[c] >= 0) {
                    if (near >= 0 && (near != r || core != h.core_of_row[c])) return no("a chain touches the core in more than one place");
                    near = r; core = h.core_of_row[c];
                }
        if (near < 0) return no("a sub-network is not connected to the block");
        Chain ch;
        ch.core = core;
        std::vector<int> order{near};
        int prev = -1, cur = near;
        for (;;) {
            int next = -1, cnt = 0;
            for (int c : adj[cur]) if (h.core_of_row[c] < 0 && c != prev) { next = c; ++cnt; }
            if (cnt > 1) return no("a sub-network hanging off the block is not a chain");
            if (cnt == 0) break;
            prev = cur; cur = next;
            order.push_back(cur);
        }
        if (order.size() != comp.size()) return no("a sub-network hanging off the block is not a chain");
        ch.rows.assign(order.rbegin(), order.rend());
        for (size_t t = 0; t < ch.rows.size(); ++t) h.pos_in_chain[ch.rows[t]] = (int)t;
        h.chains.push_back(ch);
    }
    for (size_t k = 0; k < h.chains.size(); ++k)
        for (size_t k2 = k + 1; k2 < h.chains.size(); ++k2)
            if (h.chains[k].core == h.chains[k2].core) return no("two chains on one block node");
    int total_rows = 0;
    for (const Chain& c : h.chains) total_rows += (int)c.rows.size();
    if (total_rows > HUB_MAX_CHAIN_ROWS) return no("too many chain rows for the register-resident form");
    // chain rows must be tridiagonal along the chain (+ the one coupling to the core)
    for (size_t k = 0; k < h.chains.size(); ++k) {
        const Chain& c = h.chains[k];
        for (size_t t = 0; t < c.rows.size(); ++t)
            for (const Entry& en : h.rows[c.rows[t]]) {
                if (en.col >= n) continue;
                if (h.core_of_row[en.col] >= 0) { if (t + 1 != c.rows.size() || h.core_of_row[en.col] != c.core) return no("chain row coupled to a second core node"); continue; }
                if (h.chain_of_row[en.col] != (int)k || std::abs(h.pos_in_chain[en.col] - (int)t) > 1) return no("chain is not tridiagonal");
            }
    }
    // ports: one source entry per column, on the port's own row
    std::vector<int> rhs_count(P, 0);
    for (int r = 0; r < n; ++r)
        for (const Entry& en : h.rows[r])
            if (en.col >= n) {
                const int q = en.col - n;
                if (q >= P || g.port_rows[3 * q] != r || en.neg) return no("right-hand side is not one source entry per port row");
                rhs_count[q]++;
            }
    for (int q = 0; q < P; ++q) {
        const int r = g.port_rows[3 * q];
        if (r < 0 || r >= n || rhs_count[q] != 1 || g.epi_port_of_row[r] != q) return no("port without a private signal row");
        if (h.chain_of_row[r] >= 0) h.chains[h.chain_of_row[r]].ports.push_back(q);
    }
    // do all traces of the block share one knot vector?  (then the basis is computed once per point)
    h.shared_knots = true;
    for (int e = 0; e < Pn * Pn; ++e) {
        const int kind = prefs[2 * (h.coop_first + e)], idx = prefs[2 * (h.coop_first + e) + 1];
        if (kind != PK_SPLINE || idx < 0 || (size_t)idx * 4 + 3 >= g.spl_trace.size()) { h.shared_knots = false; break; }
        const int32_t* t0 = &g.spl_trace[(size_t)prefs[2 * h.coop_first + 1] * 4];
        const int32_t* t = &g.spl_trace[(size_t)idx * 4];
        if (t[0] != t0[0] || t[1] != t0[1] || t[3] != t0[3]) { h.shared_knots = false; break; }
    }
    h.ok = true;
    return h;
}

struct HubGen {
    const GenPlan& g;
    const Hub& h;
    const int32_t* prefs;
    int LD;                                        // columns of the shared-memory tile: Pn (the matrix alone)
    long long flops = 0;
    std::string decl;                              // value declarations + code, emitted once per point

    HubGen(const GenPlan& g_, const Hub& h_, const int32_t* p) : g(g_), h(h_), prefs(p) { LD = h.Pn; }

    std::string sm(int i, int c) const { return fmt("gsm[%d * GEN_THREADS + threadIdx.x]", i * LD + c); }

    std::string cst(int idx) const {                   // baked in when the run's constants are known
        if ((size_t)idx * 2 + 1 < g.run_consts.size()) {
            const double re = g.run_consts[2 * (size_t)idx], im = g.run_consts[2 * (size_t)idx + 1];
            return "cmk(" + dlit(re) + ", " + (im == 0.0 ? std::string("0.0") : dlit(im)) + ")";
        }
        return fmt("D.pv.consts[%d]", idx);
    }

    std::string param(int pi) const {
        const int kind = prefs[2 * pi], idx = prefs[2 * pi + 1];
        switch (kind) {
            case PK_CONST: return cst(idx);
            case PK_SAMPLE: return fmt("cmk(srow[%d], 0.0)", idx);
            case PK_SAMPLE_NEG: return fmt("cmk(-srow[%d], 0.0)", idx);
            case PK_SAMPLE_INV: return fmt("cmk(1.0 / srow[%d], 0.0)", idx);
            case PK_TABLE: return fmt("D.pv.tables[%dLL * D.pv.table_ld + fi]", idx);
            case PK_SPLINE: return fmt("spline_eval(D.pv.spl_t, D.pv.spl_c, D.pv.spl_trace, D.pv.spl_fill, D.pv.spl_range, %d, f)", idx);
            case PK_SMD: return fmt("eval_smd(D.pv.consts + %d, f)", idx);
            case PK_BETA_MUL: return fmt("cmk(__dmul_rn(w, D.pv.consts[%d].x) / HVB_C0, 0.0)", idx);
            case PK_BETA: return "gen_beta(wc0, " + cst(idx) + ")";
        }
        return "cmk(0.0, 0.0)";
    }

    // every fast / simple value op as a named variable (few dozen values: they all fit in registers)
    std::string values() const {
        std::string t;
        for (int o = 0; o < g.nfast; ++o) {
            const int code = g.fop_code[o];
            std::string arg = (code & 4) ? fmt("srow[%d]", (int)g.fop_arg[o]) : dlit(g.fop_arg[o]);
            if (code & 2) arg = "__dmul_rn(w, " + arg + ")";
            if (code & 1) arg = ((code & 2) ? "-rcp_fast(" : "rcp_fast(") + arg + ")";
            t += fmt("    const double vf%d = ", o) + arg + ";\n";
        }
        for (int o = 0; o < g.nops; ++o) {
            const int32_t* op = &g.ops[(size_t)o * HVB_OP_WORDS];
            const int code = op[0];
            if (code == OP_TLINE) {
                t += fmt("    cplx vo%d[9]; eval_tline_fast(%s, %s, %s.x, vo%d);\n", o, param(op[2]).c_str(), param(op[3]).c_str(), param(op[4]).c_str(), o);
                continue;
            }
            const std::string p0 = param(op[2]);
            if (code == OP_COPY) t += fmt("    const cplx vo%d = %s;\n", o, p0.c_str());
            else if (code == OP_RECIP) t += fmt("    const cplx vo%d = crecip_np(%s);\n", o, p0.c_str());
            else if (code == OP_CAP) t += fmt("    const cplx vo%d = gen_cap(w, %s);\n", o, p0.c_str());
            else t += fmt("    const cplx vo%d = gen_ind(w, %s);\n", o, p0.c_str());
        }
        return t;
    }

    // (re, im) text of a non-block value; "0" = exactly zero
    void value(int vid, std::string* re, std::string* im) const {
        if (vid < g.nvc) {
            const double a = g.const_vals[2 * (size_t)vid], b = g.const_vals[2 * (size_t)vid + 1];
            *re = a == 0.0 ? "0" : dlit(a);
            *im = b == 0.0 ? "0" : dlit(b);
            return;
        }
        const Prod& p = h.prod[vid - g.nvc];
        if (p.kind == 0) {
            const std::string nm = fmt("vf%d", p.idx);
            if (g.fop_code[p.idx] & 2) { *re = "0"; *im = nm; } else { *re = nm; *im = "0"; }
            return;
        }
        const bool line = g.ops[(size_t)p.idx * HVB_OP_WORDS] == OP_TLINE;
        const std::string base = line ? fmt("vo%d[%d]", p.idx, p.sub) : fmt("vo%d", p.idx);
        *re = base + ".x"; *im = base + ".y";
    }

    // cell (r, col) without the block's own value: "cmk(re, im)", or "" when the cell holds nothing else
    std::string cell(int r, int col) const {
        std::string re, im;
        bool any = false;
        auto add = [](std::string& acc, const std::string& term, bool neg) {
            if (term == "0") return;
            if (acc.empty()) acc = neg ? "-" + term : term;
            else acc += (neg ? " - " : " + ") + term;
        };
        for (const Entry& en : h.rows[r]) {
            if (en.col != col) continue;
            if (en.vid >= g.nvc && h.prod[en.vid - g.nvc].kind == 2) continue;
            std::string a, b;
            value(en.vid, &a, &b);
            add(re, a, en.neg);
            add(im, b, en.neg);
            any = true;
        }
        if (!any) return "";
        return "cmk(" + (re.empty() ? std::string("0.0") : re) + ", " + (im.empty() ? std::string("0.0") : im) + ")";
    }
    std::string cell0(int r, int col) const { const std::string c = cell(r, col); return c.empty() ? "cmk(0.0, 0.0)" : c; }
};

}  // namespace

bool hvb_gen_hub_supported(const GenPlan& g, const int32_t* prefs, std::string* why) {
    if (!prefs && g.nprefs) { if (why) *why = "prefs missing"; return false; }
    const Hub h = analyse_hub(g, prefs);
    if (why) *why = h.why;
    return h.ok;
}

// structure only (no parameter kinds needed): can this plan be a hub at all?
bool hvb_gen_hub_candidate(const GenPlan& g) {
    if (g.ncoops != 1 || !g.epi_simple || g.ncols - g.n != g.P || g.P < 1 || g.P > HUB_MAX_CORE) return false;
    std::vector<int32_t> fake((size_t)std::max(g.nprefs, 1) * 2, 0);
    return analyse_hub(g, fake.data()).ok;
}

int hvb_gen_hub_source(const GenPlan& g, const int32_t* prefs, const char* prelude, GenResult* out, std::string* err) {
    const Hub h = analyse_hub(g, prefs);
    if (!h.ok) { if (err) *err = h.why; return -1; }
    HubGen G(g, h, prefs);
    const int n = g.n, P = g.P, Pn = h.Pn, LD = G.LD;
    (void)n;
    const int threads = 32;
    const size_t smem = (size_t)Pn * LD * threads * 16;
    std::string b;                                 // body of one point

    // ---- 1. S_ij(f) -> tile [(I+S)^T | (I-S)^T], row j column i <- S_ij ------------------------------------------
    b += "    // ---- N-port block: S(f) -> Y = (1/Z0)(I - S)(I + S)^-1   (base/nport.py:28-39)\n";
    if (h.shared_knots) {
        const int32_t* t0 = &g.spl_trace[(size_t)prefs[2 * h.coop_first + 1] * 4];
        b += "    { SplineBasis sb;\n";
        b += fmt("      spline_basis(D.pv.spl_t, %d, %d, %d, f, &sb);\n", t0[0], t0[1], t0[3]);
        for (int i = 0; i < Pn; ++i)
            for (int j = 0; j < Pn; ++j) {
                const int tr = prefs[2 * (h.coop_first + i * Pn + j) + 1];
                const int32_t* t = &g.spl_trace[(size_t)tr * 4];
                b += fmt("      { const double2 rg = D.pv.spl_range[%d]; const cplx s = f < rg.x ? D.pv.spl_fill[%d] : (f > rg.y ? D.pv.spl_fill[%d] : spline_dot(D.pv.spl_c + %d, sb));\n",
                         tr, 2 * tr, 2 * tr + 1, t[2]);
                b += fmt("        %s = cmk(%s + s.x, s.y); }\n", G.sm(i, j).c_str(), i == j ? "1.0" : "0.0");
                G.flops += 16;
            }
        b += "    }\n";
    } else {
        for (int i = 0; i < Pn; ++i)
            for (int j = 0; j < Pn; ++j) {
                b += fmt("    { const cplx s = %s;\n", G.param(h.coop_first + i * Pn + j).c_str());
                b += fmt("      %s = cmk(%s + s.x, s.y); }\n", G.sm(i, j).c_str(), i == j ? "1.0" : "0.0");
            }
    }
    b += fmt("    gen_dense_inv<%d, %d>(gsm + threadIdx.x, pmin, pmax);              // the tile now holds W = (I + S)^-1\n", Pn, LD);
    G.flops += 8LL * Pn * Pn * Pn + 8LL * Pn * Pn;
    // core matrix A[i][j] = Y[i][j] = (2 W[i][j] - delta_ij) / Z0 (+ whatever else is stamped on core cells), in place
    b += "    { const cplx iz = crecip_np(" + G.cst(h.coop_z0) + ");\n";
    for (int i = 0; i < Pn; ++i)
        for (int j = 0; j < Pn; ++j) {
            b += fmt("      { cplx x = %s; x.x = fma(2.0, x.x, %s); x.y = 2.0 * x.y; cplx y = (iz.y == 0.0) ? cscale(x, iz.x) : cmul(iz, x);",
                     G.sm(i, j).c_str(), i == j ? "-1.0" : "0.0");
            const std::string extra = G.cell(h.core_row[i], h.core_row[j]);
            if (!extra.empty()) b += " const cplx e = " + extra + "; y.x += e.x; y.y += e.y;";
            b += " " + G.sm(i, j) + " = y; }\n";
        }
    b += "    }\n";
    // right-hand sides of the core system: port q's column has ONE entry, at core node cq[q] (its own node, or its chain's)
    std::string atdecl;
    std::vector<int> cq(P, -1);
    for (int q = 0; q < P; ++q) {
        const int r = g.port_rows[3 * q];
        if (h.core_of_row[r] >= 0) {
            cq[q] = h.core_of_row[r];
            b += fmt("    const cplx rb%d = %s;\n", q, G.cell0(r, g.n + q).c_str());
        } else {
            cq[q] = h.chains[h.chain_of_row[r]].core;
            atdecl += fmt("    cplx rb%d = cmk(0.0, 0.0);\n", q);
        }
    }

    // ---- 2. chains: front-to-back elimination in registers, folded into the core node -------------------------------
    std::string chdecl;
    for (size_t k = 0; k < h.chains.size(); ++k) {
        const Chain& c = h.chains[k];
        const int M = (int)c.rows.size(), ci = c.core, crow = h.core_row[ci];
        const int nq = (int)c.ports.size();
        b += fmt("    // ---- chain %d: %d row(s) hanging off block port %d, %d port(s)\n", (int)k, M, ci, nq);
        // state: cur row (c0 c1 c2) at columns t, t+1, t+2 of the chain (column M = the core node), its right-hand sides
        b += "    {\n      cplx c0, c1, c2 = cmk(0.0, 0.0), n0, n1, n2;\n";
        for (int qi = 0; qi < nq; ++qi) b += fmt("      cplx yc%d, yn%d;\n", qi, qi);
        auto col_of = [&](int t) { return t < M ? c.rows[t] : crow; };
        auto rhs_init = [&](const char* pre, int t) {
            std::string s;
            for (int qi = 0; qi < nq; ++qi) {
                const int q = c.ports[qi];
                s += fmt("      %s%d = %s;\n", pre, qi, (g.port_rows[3 * q] == c.rows[t]) ? G.cell0(c.rows[t], g.n + q).c_str() : "cmk(0.0, 0.0)");
            }
            return s;
        };
        b += "      c0 = " + G.cell0(c.rows[0], col_of(0)) + "; c1 = " + G.cell0(c.rows[0], col_of(1)) + ";\n";
        b += rhs_init("yc", 0);
        for (int t = 0; t < M; ++t) {
            const bool last = (t == M - 1);
            // stored pivot row t of chain k: inverse pivot, two super-diagonal entries, right-hand sides
            chdecl += fmt("    cplx ui%d_%d, ua%d_%d, ub%d_%d;\n", (int)k, t, (int)k, t, (int)k, t);
            for (int qi = 0; qi < nq; ++qi) chdecl += fmt("    cplx uy%d_%d_%d;\n", (int)k, t, qi);
            if (!last) {
                b += "      n0 = " + G.cell0(c.rows[t + 1], col_of(t)) + "; n1 = " + G.cell0(c.rows[t + 1], col_of(t + 1)) + "; n2 = " +
                     (t + 2 <= M ? G.cell0(c.rows[t + 1], col_of(t + 2)) : std::string("cmk(0.0, 0.0)")) + ";\n";
                b += rhs_init("yn", t + 1);
                b += "      { double best = fabs(c0.x) + fabs(c0.y); if (!(best >= 0.0)) best = -1.0;\n";
                b += "        const double mn = fabs(n0.x) + fabs(n0.y); const bool sw = mn > best; if (sw) best = mn;\n";
                b += "        pmin = fmin(pmin, best); pmax = fmax(pmax, best);\n";
                b += "        const cplx p0 = sw ? n0 : c0, p1 = sw ? n1 : c1, p2 = sw ? n2 : c2;\n";
                b += "        const cplx r0 = sw ? c0 : n0; cplx r1 = sw ? c1 : n1, r2 = sw ? c2 : n2;\n";
                b += "        const cplx inv = crecip_fast(p0); const cplx m = cmul(r0, inv); cfnma(r1, m, p1); cfnma(r2, m, p2);\n";
                for (int qi = 0; qi < nq; ++qi) {
                    b += fmt("        { const cplx py = sw ? yn%d : yc%d; cplx ry = sw ? yc%d : yn%d; cfnma(ry, m, py); uy%d_%d_%d = py; yc%d = ry; }\n",
                             qi, qi, qi, qi, (int)k, t, qi, qi);
                    G.flops += 8;
                }
                b += fmt("        ui%d_%d = inv; ua%d_%d = p1; ub%d_%d = p2; c0 = r1; c1 = r2; c2 = cmk(0.0, 0.0); }\n", (int)k, t, (int)k, t, (int)k, t);
                G.flops += 8 + 6 + 16;
            } else {
                // last chain row against its core row: no exchange, growth check instead (see the header comment)
                const std::string e = G.cell0(crow, c.rows[t]);
                b += "      { double best = fabs(c0.x) + fabs(c0.y); if (!(best >= 0.0)) best = -1.0;\n";
                b += "        pmin = fmin(pmin, best); pmax = fmax(pmax, best);\n";
                b += "        const cplx ec = " + e + ";\n";
                b += "        if ((fabs(ec.x) + fabs(ec.y)) > 1e4 * best) bad |= HVB_STATUS_ZERO_PIVOT_BIT;\n";
                b += "        const cplx inv = crecip_fast(c0); const cplx m = cmul(ec, inv);\n";
                b += fmt("        { cplx d = %s; cfnma(d, m, c1); %s = d; }\n", G.sm(ci, ci).c_str(), G.sm(ci, ci).c_str());
                for (int qi = 0; qi < nq; ++qi) {
                    const int q = c.ports[qi];
                    b += fmt("        cfnma(rb%d, m, yc%d); uy%d_%d_%d = yc%d;\n", q, qi, (int)k, t, qi, qi);
                    G.flops += 8;
                }
                b += fmt("        ui%d_%d = inv; ua%d_%d = c1; ub%d_%d = cmk(0.0, 0.0); }\n", (int)k, t, (int)k, t, (int)k, t);
                G.flops += 8 + 6 + 8;
            }
        }
        b += "    }\n";
    }

    // ---- 3. core solve ----------------------------------------------------------------------------------------------------
    b += fmt("    gen_dense_inv<%d, %d>(gsm + threadIdx.x, pmin, pmax);              // the tile now holds Z = (core matrix)^-1\n", Pn, LD);
    G.flops += 8LL * Pn * Pn * Pn + 8LL * Pn * Pn;
    auto xcore = [&](int i, int q) { return fmt("cmul(%s, rb%d)", G.sm(i, cq[q]).c_str(), q); };   // core voltage i of excitation q

    // ---- 4. back substitution along the chains, S = (2 V_sig - delta) sqrt|Zi| / sqrt|Zj| ---------------------------------
    auto emit_S = [&](int pj, int i, const std::string& x) {
        std::string t = "        { cplx sv = " + x + fmt("; sv.x = fma(2.0, sv.x, %s); sv.y = 2.0 * sv.y; ", pj == i ? "-1.0" : "0.0");
        const double sc = g.epi_scale[(size_t)pj * P + i];
        if (sc != 1.0) t += fmt("sv = cscale(sv, %a); ", sc);
        t += "if (!(isfinite(sv.x) && isfinite(sv.y))) bad |= HVB_STATUS_NONFINITE_BIT; ";
        t += fmt("Sp[%dLL * D.ldS] = sv; }\n", pj * P + i);
        return t;
    };
    b += "    // ---- port voltages -> S (solvers/spfn.py:159-176, closed form for ground-referenced ports)\n";
    for (int q = 0; q < P; ++q) {
        const int r = g.port_rows[3 * q];
        if (h.core_of_row[r] >= 0)
            for (int i = 0; i < P; ++i) { b += emit_S(q, i, xcore(h.core_of_row[r], i)); G.flops += 6; }
    }
    for (size_t k = 0; k < h.chains.size(); ++k) {
        const Chain& c = h.chains[k];
        if (c.ports.empty()) continue;
        const int M = (int)c.rows.size();
        int first = M;                               // farthest row a port sits on
        for (int q : c.ports) first = std::min(first, h.pos_in_chain[g.port_rows[3 * q]]);
        for (int i = 0; i < P; ++i) {                 // excitation i
            b += fmt("    { cplx x1 = %s, x2 = cmk(0.0, 0.0);\n", xcore(c.core, i).c_str());
            G.flops += 6;
            for (int t = M - 1; t >= first; --t) {
                int qi_of_i = -1;
                for (size_t qi = 0; qi < c.ports.size(); ++qi) if (c.ports[qi] == i) qi_of_i = (int)qi;
                b += "      { cplx x = " + (qi_of_i >= 0 ? fmt("uy%d_%d_%d", (int)k, t, qi_of_i) : std::string("cmk(0.0, 0.0)")) + ";";
                b += fmt(" cfnma(x, ua%d_%d, x1); cfnma(x, ub%d_%d, x2); x = cmul(x, ui%d_%d); x2 = x1; x1 = x;\n", (int)k, t, (int)k, t, (int)k, t);
                G.flops += 8 + 8 + 6;
                for (int q : c.ports)
                    if (h.pos_in_chain[g.port_rows[3 * q]] == t) b += emit_S(q, i, "x");
                b += "      }\n";
            }
            b += "    }\n";
        }
    }

    std::string src;
    src += "// generated by heavi_b200 (csrc/hvb_gen_hub.cpp): kernel E for one netlist (N-port block + matching chains)\n";
    src += prelude;
    src += fmt("\n#define GEN_THREADS %d\n", threads);
    src += "__device__ __forceinline__ cplx gen_beta(double b0, cplx s) {          // b0 = (2 pi f) / c0, once per point\n"
           "    return cmk(__dmul_rn(b0, s.x), s.y == 0.0 ? 0.0 : __dmul_rn(b0, s.y));\n}\n"
           "__device__ __forceinline__ cplx gen_cap(double w, cplx p0) {\n"
           "    return (p0.y == 0.0) ? cmk(0.0, __dmul_rn(w, p0.x)) : cmk(-__dmul_rn(w, p0.y), __dmul_rn(w, p0.x));\n}\n"
           "__device__ __forceinline__ cplx gen_ind(double w, cplx p0) {\n"
           "    if (p0.y == 0.0) return cmk(0.0, -rcp_fast(__dmul_rn(w, p0.x)));\n"
           "    return cdiv_np(cmk(1.0, 0.0), cmk(-__dmul_rn(w, p0.y), __dmul_rn(w, p0.x)));\n}\n"
           "// In-place inverse of a per-thread NR x NR tile in shared memory (row stride LD, element e of this thread at\n"
           "// t[e * GEN_THREADS]): Gauss-Jordan with partial pivoting.  Rows are exchanged physically while eliminating; the\n"
           "// exchanges are undone on the columns afterwards, last first (the inverse of P A is A^-1 P^-1).  Every index is a\n"
           "// loop counter or one dynamic row / column: no per-thread index arrays in local memory.\n"
           "template <int NR, int LD>\n"
           "__device__ __forceinline__ void gen_dense_inv(cplx* __restrict__ t, double& pmin, double& pmax) {\n"
           "#define TE(i, c) t[((i) * LD + (c)) * GEN_THREADS]\n"
           "    int perm[NR];                                          // only ever indexed by unrolled loop counters: registers\n"
           "#pragma unroll\n"
           "    for (int k = 0; k < NR; ++k) {\n"
           "        int p = k;\n"
           "        double best = fabs(TE(k, k).x) + fabs(TE(k, k).y);\n"
           "        if (!(best >= 0.0)) best = -1.0;\n"
           "#pragma unroll\n"
           "        for (int i = k + 1; i < NR; ++i) {\n"
           "            const cplx v = TE(i, k);\n"
           "            const double m = fabs(v.x) + fabs(v.y);\n"
           "            if (m > best) { best = m; p = i; }\n"
           "        }\n"
           "        pmin = fmin(pmin, best); pmax = fmax(pmax, best);\n"
           "        perm[k] = p;\n"
           "        if (p != k) {\n"
           "#pragma unroll\n"
           "            for (int c = 0; c < NR; ++c) { const cplx a = TE(k, c); TE(k, c) = TE(p, c); TE(p, c) = a; }\n"
           "        }\n"
           "        const cplx inv = crecip_fast(TE(k, k));\n"
           "        cplx pr[NR];                                       // the scaled pivot row stays in registers for this step\n"
           "#pragma unroll\n"
           "        for (int c = 0; c < NR; ++c) { pr[c] = (c == k) ? inv : cmul(TE(k, c), inv); TE(k, c) = pr[c]; }\n"
           "#pragma unroll 1\n"
           "        for (int i = 0; i < NR; ++i) {                     // rows stay a loop: bounded register use and code size\n"
           "            if (i == k) continue;\n"
           "            cplx* __restrict__ rowp = t + i * (LD * GEN_THREADS);\n"
           "            const cplx m = rowp[k * GEN_THREADS];\n"
           "#pragma unroll\n"
           "            for (int c = 0; c < NR; ++c) { cplx v = (c == k) ? cmk(0.0, 0.0) : rowp[c * GEN_THREADS]; cfnma(v, m, pr[c]); rowp[c * GEN_THREADS] = v; }\n"
           "        }\n"
           "    }\n"
           "#pragma unroll\n"
           "    for (int k = NR - 1; k >= 0; --k) {\n"
           "        const int p = perm[k];\n"
           "        if (p != k) {\n"
           "#pragma unroll\n"
           "            for (int i = 0; i < NR; ++i) { const cplx a = TE(i, k); TE(i, k) = TE(i, p); TE(i, p) = a; }\n"
           "        }\n"
           "    }\n"
           "#undef TE\n}\n\n";
    out->kernel_name = "hvb_gen_kernel";
    src += "extern \"C\" __global__ void __launch_bounds__(GEN_THREADS) hvb_gen_kernel(const HvbDev D) {\n";
    src += "#ifdef HVB_HOST_SIM\n";
    src += fmt("  static cplx gsm[%d * GEN_THREADS];\n", Pn * LD);
    src += "#else\n  extern __shared__ cplx gsm[];\n#endif\n";
    src += "  const long long T = (long long)gridDim.x * blockDim.x;\n";
    src += "  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n";
    src += "  const long long total = D.nS * D.nF;\n  unsigned bad = 0;\n";
    src += "  for (long long q = t0; q < total; q += T) {\n";
    src += "    long long s = 0, fi = q;\n";
    src += "    if (D.nS != 1) { s = q / D.nF; fi = q - s * D.nF; }\n";
    src += "    const double f = __ldg(D.f + fi);\n";
    src += "    if (D.f32 && s == 0) D.f32[fi] = (float)f;\n";
    src += "    const double* __restrict__ srow = D.samples + s * D.nR; (void)srow;\n";
    src += "    const double w = __dmul_rn(HVB_TWOPI, f); (void)w;\n";
    src += "    const double wc0 = w / HVB_C0; (void)wc0;\n";
    src += fmt("    cplx* __restrict__ Sp = D.S + s * %dLL * D.ldS + fi;\n", P * P);
    src += "    double pmin = 1e300, pmax = 0.0;\n";
    src += G.values();
    src += atdecl + chdecl + b;
    src += "    if (!(pmin >= 2.2250738585072014e-308)) bad |= HVB_STATUS_ZERO_PIVOT_BIT;\n";
    src += "    if (!(pmax <= 1.7976931348623157e308)) bad |= HVB_STATUS_NONFINITE_BIT;\n";
    src += "  }\n  if (bad && D.status) atomicOr(D.status, (int)bad);\n}\n";
    out->src.swap(src);
    out->rolled = false; out->scratch_acc = false; out->tables_in_source = false; out->n_cases = 0;
    out->rows.clear(); out->itab.clear(); out->dtab.clear();
    out->threads = threads;
    out->smem_bytes = smem;
    out->flops_per_solve = G.flops;
    return 0;
}
