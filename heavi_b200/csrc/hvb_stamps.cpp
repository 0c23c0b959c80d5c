// hvb_stamps.cpp -- the stamp-table -> device-plan compiler below the C-ABI (hvb_plan_create_from_stamps).
//
// What the reference does per run_sp call -- walk the components, let each add its (k,k,nF) block into the dense
// tensor at the MNA indices of its `ids` (network.py:385-404, base/*.py) -- is frozen here ONCE into the arrays of
// hvb_plan_desc: a value program (which stamp values depend on frequency / sample and how), a cell program (which
// values sum into which matrix cell, in component order, with numpy's last-write-wins rule for a component that
// repeats an index), the port tables of the S epilogue, and a row order.  Same algorithm, same
// output arrays as heavi_b200/plan.py:compile_plan (tests/test_stamps.py compares them entry by entry), so any host
// language that can fill an hvb_stamp_table gets the whole path without Python.
#include "hvb_stamps.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <queue>
#include <set>

#include "plan_format.h"

namespace {

const long long DYN = 1LL << 28;        // marker added to per-point value ids until the constant count is known

struct cd { double re, im; };

// numpy's complex division (Smith's algorithm with a reciprocal scale), for 1 / z folded at plan time
cd cdiv_np(cd a, cd b) {
    const double abr = std::fabs(b.re), abi = std::fabs(b.im);
    if (abr >= abi) {
        if (abr == 0.0 && abi == 0.0) return cd{a.re / abr, a.im / abi};
        const double rat = b.im / b.re, scl = 1.0 / (b.re + b.im * rat);
        return cd{(a.re + a.im * rat) * scl, (a.im - a.re * rat) * scl};
    }
    const double rat = b.re / b.im, scl = 1.0 / (b.im + b.re * rat);
    return cd{(a.re * rat + a.im) * scl, (a.im * rat - a.re) * scl};
}

struct Builder {
    std::vector<cd> consts;
    std::vector<int32_t> prefs;                 // (kind, idx) pairs
    std::vector<cd> const_vals{cd{1.0, 0.0}};   // value id 0 == 1
    int dyn_count = 0;
    std::vector<int32_t> ops, coops;
    int coop_scratch = 0, n_tables = 0, n_sample_cols = 0, max_nport = 0;
    std::vector<int32_t> table_pref, table_slot;   // per table parameter, in order of appearance

    int cst(cd v) { consts.push_back(v); return (int)consts.size() - 1; }
    int pref(int kind, int idx) { prefs.push_back(kind); prefs.push_back(idx); return (int)prefs.size() / 2 - 1; }
    int const_value(cd v) { const_vals.push_back(v); return (int)const_vals.size() - 1; }
    int dyn_values(int count) { const int b = dyn_count; dyn_count += count; return b; }

    // -> pref index; *is_const / *value describe a parameter fixed for the whole run (plan.py _Builder.param)
    int param(const hvb_param& p, bool* is_const, cd* value) {
        *is_const = false;
        switch (p.kind) {
            case HVB_PAR_CONST:
                *is_const = true; *value = cd{p.v[0], p.v[1]};
                return pref(PK_CONST, cst(*value));
            case HVB_PAR_SAMPLE: n_sample_cols = std::max(n_sample_cols, p.index + 1); return pref(PK_SAMPLE, p.index);
            case HVB_PAR_SAMPLE_NEG: n_sample_cols = std::max(n_sample_cols, p.index + 1); return pref(PK_SAMPLE_NEG, p.index);
            case HVB_PAR_SAMPLE_INV: n_sample_cols = std::max(n_sample_cols, p.index + 1); return pref(PK_SAMPLE_INV, p.index);
            case HVB_PAR_BETA: return pref(PK_BETA, cst(cd{p.v[0], p.v[1]}));           // v = sqrt(er), taken by the host
            case HVB_PAR_SPLINE: return pref(PK_SPLINE, p.index);
            case HVB_PAR_BETA_MUL: return pref(PK_BETA_MUL, cst(cd{p.v[0], 0.0}));
            case HVB_PAR_SMD: {
                const int first = cst(cd{(double)p.index, 0.0});
                cst(cd{p.v[0], 0.0}); cst(cd{p.v[1], 0.0}); cst(cd{p.v[2], 0.0});
                return pref(PK_SMD, first);
            }
            default: {                                                                   // HVB_PAR_TABLE
                const int slot = cst(cd{0.0, 0.0});                                      // the slot a scalar-valued callable falls back to
                n_tables = std::max(n_tables, p.index + 1);
                const int pi = pref(PK_TABLE, p.index);
                table_pref.push_back(pi); table_slot.push_back(slot);
                return pi;
            }
        }
    }
};

struct Block { int k = 0; std::vector<long long> vid; std::vector<int8_t> neg; int zpref = -1; };

int component_values(Builder& b, int kind, const hvb_param* par, int npar, int nids, Block* out, std::string* err) {
    auto block2 = [&](long long vid) {
        out->k = 2; out->vid = {vid, vid, vid, vid}; out->neg = {0, 1, 1, 0};
    };
    bool st; cd v;
    if (kind == HVB_COMP_PORT) {
        if (npar < 1 || nids != 4) { *err = "port needs 1 parameter and 4 ids"; return -1; }
        const int pi = b.param(par[0], &st, &v);
        long long g;
        if (st) g = b.const_value(v.im == 0.0 ? cd{1.0 / v.re, 0.0} : cdiv_np(cd{1.0, 0.0}, v));
        else {
            const int d = b.dyn_values(1);
            g = DYN + d;
            const int32_t op[HVB_OP_WORDS] = {OP_RECIP, d, pi, 0, 0, 0, 0, 0};
            b.ops.insert(b.ops.end(), op, op + HVB_OP_WORDS);
        }
        out->k = 4; out->vid.assign(16, -1); out->neg.assign(16, 0);
        const int pat[8][4] = {{0, 0, 1, 0}, {0, 1, 1, 1}, {1, 0, 1, 1}, {1, 1, 1, 0}, {1, 3, 0, 0}, {2, 3, 0, 1}, {3, 1, 0, 0}, {3, 2, 0, 1}};
        for (const auto& e : pat) { out->vid[e[0] * 4 + e[1]] = e[2] ? g : 0; out->neg[e[0] * 4 + e[1]] = (int8_t)e[3]; }
        out->zpref = pi;
        return 0;
    }
    if (kind == HVB_COMP_RESISTOR) {
        if (npar < 1 || par[0].kind != HVB_PAR_CONST) { *err = "resistor needs a constant conductance"; return -1; }
        block2(b.const_value(cd{par[0].v[0], par[0].v[1]}));
        return 0;
    }
    if (kind == HVB_COMP_ADMITTANCE || kind == HVB_COMP_IMPEDANCE || kind == HVB_COMP_CAPACITOR || kind == HVB_COMP_INDUCTOR) {
        if (npar < 1) { *err = "two-terminal part without a parameter"; return -1; }
        const int pi = b.param(par[0], &st, &v);
        if (st && kind == HVB_COMP_ADMITTANCE) { block2(b.const_value(v)); return 0; }
        if (st && kind == HVB_COMP_IMPEDANCE) { block2(b.const_value(v.im == 0.0 ? cd{1.0 / v.re, 0.0} : cdiv_np(cd{1.0, 0.0}, v))); return 0; }
        const int opcode = kind == HVB_COMP_ADMITTANCE ? OP_COPY : kind == HVB_COMP_IMPEDANCE ? OP_RECIP : kind == HVB_COMP_CAPACITOR ? OP_CAP : OP_IND;
        const int d = b.dyn_values(1);
        const int32_t op[HVB_OP_WORDS] = {opcode, d, pi, 0, 0, 0, 0, 0};
        b.ops.insert(b.ops.end(), op, op + HVB_OP_WORDS);
        block2(DYN + d);
        return 0;
    }
    if (kind == HVB_COMP_TLINE) {
        if (npar < 3 || nids != 3) { *err = "line needs Z0, beta, length and 3 ids"; return -1; }
        const int pz = b.param(par[0], &st, &v), pb = b.param(par[1], &st, &v), plen = b.param(par[2], &st, &v);
        const int base = b.dyn_values(9);
        const int32_t op[HVB_OP_WORDS] = {OP_TLINE, base, pz, pb, plen, 0, 0, 0};
        b.ops.insert(b.ops.end(), op, op + HVB_OP_WORDS);
        const long long t = DYN + base;
        out->k = 3; out->vid = {t + 0, t + 1, t + 4, t + 2, t + 3, t + 5, t + 6, t + 7, t + 8}; out->neg.assign(9, 0);
        return 0;
    }
    if (kind == HVB_COMP_NPORT_S || kind == HVB_COMP_NPORT_Y) {
        const int Pn = nids - 1;
        if (Pn < 1 || Pn > HVB_MAX_NPORT) { *err = "N-port block with an unsupported number of ports"; return -1; }
        if (npar < Pn * Pn + (kind == HVB_COMP_NPORT_S ? 1 : 0)) { *err = "N-port block: missing parameters"; return -1; }
        const int first = (int)b.prefs.size() / 2;
        for (int e = 0; e < Pn * Pn; ++e) b.param(par[e], &st, &v);
        const int base = b.dyn_values((Pn + 1) * (Pn + 1));
        if (kind == HVB_COMP_NPORT_S) {
            const int z0 = b.cst(cd{par[Pn * Pn].v[0], par[Pn * Pn].v[1]});
            const int32_t co[HVB_COOP_WORDS] = {COOP_NPORT_S, base, Pn, first, z0, 0, 0, 0};
            b.coops.insert(b.coops.end(), co, co + HVB_COOP_WORDS);
            b.coop_scratch = std::max(b.coop_scratch, 2 * Pn * Pn);
        } else {
            const int32_t co[HVB_COOP_WORDS] = {COOP_NPORT_Y, base, Pn, first, 0, 0, 0, 0};
            b.coops.insert(b.coops.end(), co, co + HVB_COOP_WORDS);
            b.coop_scratch = std::max(b.coop_scratch, Pn * Pn);
        }
        b.max_nport = std::max(b.max_nport, Pn);
        out->k = Pn + 1; out->vid.resize((size_t)(Pn + 1) * (Pn + 1)); out->neg.assign((size_t)(Pn + 1) * (Pn + 1), 0);
        for (int e = 0; e < (Pn + 1) * (Pn + 1); ++e) out->vid[e] = DYN + base + e;
        return 0;
    }
    *err = "unknown component kind";
    return -1;
}

bool norton_legal(const hvb_stamp_table& t) {
    if (t.P == 0) return false;
    std::map<int, std::set<int>> used;
    for (int ci = 0; ci < t.n_components; ++ci)
        for (int e = t.ids_off[ci]; e < t.ids_off[ci + 1]; ++e) used[t.ids[e]].insert(ci);
    std::set<int> seen_int;
    for (int ci = 0; ci < t.n_components; ++ci) {
        if (t.comp_kind[ci] != HVB_COMP_PORT) continue;
        const int32_t* id = t.ids + t.ids_off[ci];
        const int sig = id[0], internal = id[1], gnd = id[2], src = id[3];
        if (internal == 0 || internal == sig || internal == gnd || seen_int.count(internal)) return false;
        if (used[internal] != std::set<int>{ci} || used[src] != std::set<int>{ci}) return false;
        seen_int.insert(internal);
    }
    return true;
}

// Reverse Cuthill-McKee on the structure of the system rows: start every component at a node of minimum degree,
// visit neighbours in order of increasing degree (ties: lower index), reverse.  Chains come out tri-diagonal.
bool band_order(const hvb_stamp_table& t, std::vector<long long>& rowmap, bool norton, bool small, int* band_out) {
    long long n = 0;
    for (long long r : rowmap) n = std::max(n, r + 1);
    std::vector<std::set<int>> adj((size_t)n);
    bool any = false;
    for (int ci = 0; ci < t.n_components; ++ci) {
        std::vector<int> ids(t.ids + t.ids_off[ci], t.ids + t.ids_off[ci + 1]);
        if (t.comp_kind[ci] == HVB_COMP_PORT && norton) ids = {ids[0], ids[2]};
        std::vector<int> rows;
        for (int i : ids) if (rowmap[i] >= 0) rows.push_back((int)rowmap[i]);
        for (int a : rows) for (int c : rows) { any = true; if (a != c) adj[a].insert(c); }
    }
    if (!any || n == 0) return false;
    std::vector<int> order;
    std::vector<char> seen((size_t)n, 0);
    for (;;) {
        int start = -1;
        for (int r = 0; r < n; ++r) if (!seen[r] && (start < 0 || adj[r].size() < adj[start].size())) start = r;
        if (start < 0) break;
        std::queue<int> q;
        q.push(start); seen[start] = 1;
        while (!q.empty()) {
            const int r = q.front(); q.pop();
            order.push_back(r);
            std::vector<int> nb;
            for (int c : adj[r]) if (!seen[c]) nb.push_back(c);
            std::sort(nb.begin(), nb.end(), [&](int a, int c) { return adj[a].size() != adj[c].size() ? adj[a].size() < adj[c].size() : a < c; });
            for (int c : nb) { seen[c] = 1; q.push(c); }
        }
    }
    std::reverse(order.begin(), order.end());
    std::vector<int> pos((size_t)n);
    for (int k = 0; k < n; ++k) pos[order[k]] = k;
    int band = 0, natural = 0;
    for (int r = 0; r < n; ++r) for (int c : adj[r]) { band = std::max(band, std::abs(pos[r] - pos[c])); natural = std::max(natural, std::abs(r - c)); }
    if (natural <= band) { for (int k = 0; k < n; ++k) pos[k] = k; band = natural; }
    if (band > 8 || (small ? 2 * band + 1 > n : 2 * (3 * band + 1) > n)) return false;
    for (long long& r : rowmap) if (r >= 0) r = pos[r];
    *band_out = band;
    return true;
}

}  // namespace

bool hvb_stamps_norton_legal(const hvb_stamp_table* t) { return norton_legal(*t); }

int hvb_stamps_compile(const hvb_stamp_table* tp, int mode, int pad_rows, int pad_rhs, int order, StampPlan* out, std::string* err) {
    const hvb_stamp_table& t = *tp;
    std::string e_;
    if (!err) err = &e_;
    if (mode == HVB_MODE_AUTO) mode = norton_legal(t) ? HVB_MODE_NORTON : HVB_MODE_FULL;
    if (mode == HVB_MODE_NORTON && !norton_legal(t)) { *err = "Norton port folding is not legal for this netlist"; return -1; }
    if (mode != HVB_MODE_FULL && mode != HVB_MODE_NORTON && mode != HVB_MODE_DUMP) { *err = "bad mode"; return -1; }
    Builder b;
    const int NM = t.N + t.M, P = t.P;
    // ---- MNA index -> system row
    std::vector<long long> rowmap((size_t)NM, ROW_GROUND);
    if (mode == HVB_MODE_DUMP) for (int i = 0; i < NM; ++i) rowmap[i] = i;
    else if (mode == HVB_MODE_FULL) for (int i = 1; i < NM; ++i) rowmap[i] = i - 1;
    else {
        std::set<int> dropped;
        for (int q = 0; q < P; ++q) { dropped.insert(t.port_src[q]); dropped.insert(t.port_table[4 * q + 2]); }
        long long k = 0;
        for (int i = 1; i < NM; ++i) if (!dropped.count(i)) rowmap[i] = k++;
    }
    long long n_real = 0;
    for (long long r : rowmap) n_real = std::max(n_real, r + 1);
    int band = -1;
    const bool padded = pad_rows > 0;
    if (mode != HVB_MODE_DUMP && !padded && (order == HVB_ORDER_RCM || (order == HVB_ORDER_AUTO && n_real >= 33))) {
        bool twoports = true;
        for (int ci = 0; ci < t.n_components; ++ci)
            if ((t.comp_kind[ci] == HVB_COMP_NPORT_S || t.comp_kind[ci] == HVB_COMP_NPORT_Y) && t.ids_off[ci + 1] - t.ids_off[ci] != 3) twoports = false;
        if (twoports) band_order(t, rowmap, mode == HVB_MODE_NORTON, order == HVB_ORDER_RCM, &band);
    }
    long long n = n_real, n_rhs = mode == HVB_MODE_DUMP ? 0 : P;
    if (padded && mode != HVB_MODE_DUMP) {
        if (pad_rows < n_real || pad_rhs < P) { *err = "pad shape is smaller than the system"; return -1; }
        n = pad_rows; n_rhs = pad_rhs;
    }
    const long long ncols = n + n_rhs;

    std::map<std::pair<int, int>, std::vector<std::pair<long long, int>>> cells;
    auto add = [&](long long r, long long c, long long vid, int neg) { cells[{(int)r, (int)c}].push_back({vid, neg}); };
    struct PortInfo { int zpref; long long g; int ids[4]; };
    std::map<int, PortInfo> port_pref;
    for (int ci = 0; ci < t.n_components; ++ci) {
        const int kind = t.comp_kind[ci];
        const int32_t* ids = t.ids + t.ids_off[ci];
        const int k = t.ids_off[ci + 1] - t.ids_off[ci];
        Block blk;
        if (component_values(b, kind, t.par + t.par_off[ci], t.par_off[ci + 1] - t.par_off[ci], k, &blk, err) != 0) return -1;
        if (kind == HVB_COMP_PORT) {
            port_pref[t.comp_port[ci]] = PortInfo{blk.zpref, blk.vid[0], {ids[0], ids[1], ids[2], ids[3]}};
            if (mode == HVB_MODE_NORTON) {
                const int sig = ids[0], gnd = ids[2];
                std::vector<std::pair<std::pair<int, int>, int>> last;          // numpy-style last write wins if sig == gnd
                auto put = [&](int mr, int mc, int s) {
                    for (auto& kv : last) if (kv.first == std::make_pair(mr, mc)) { kv.second = s; return; }
                    last.push_back({{mr, mc}, s});
                };
                put(sig, sig, 0); put(sig, gnd, 1); put(gnd, sig, 1); put(gnd, gnd, 0);
                for (auto& kv : last)
                    if (rowmap[kv.first.first] >= 0 && rowmap[kv.first.second] >= 0) add(rowmap[kv.first.first], rowmap[kv.first.second], blk.vid[0], kv.second);
                continue;
            }
        }
        if (blk.k != k) { *err = "component ids do not match its kind"; return -1; }
        // last (row-major) block position writing each distinct cell wins (numpy `+=` with repeated indices)
        std::map<std::pair<int, int>, std::pair<int, int>> last;
        for (int a = 0; a < k; ++a) for (int c = 0; c < k; ++c) last[{ids[a], ids[c]}] = {a, c};
        for (auto& kv : last) {
            const int a = kv.second.first, c = kv.second.second;
            if (blk.vid[(size_t)a * k + c] < 0) continue;
            if (rowmap[kv.first.first] < 0 || rowmap[kv.first.second] < 0) continue;
            add(rowmap[kv.first.first], rowmap[kv.first.second], blk.vid[(size_t)a * k + c], blk.neg[(size_t)a * k + c]);
        }
    }
    // ---- right-hand sides and S epilogue tables
    out->port_rows.assign((size_t)P * 3, ROW_GROUND);
    out->port_zpref.assign((size_t)std::max(P, 0), 0);
    for (int q = 0; q < P; ++q) {
        const int iport = t.port_table[4 * q] + 1;
        auto it = port_pref.find(iport);
        if (it == port_pref.end()) { *err = "port table names a port that is not a component"; return -1; }
        const PortInfo& pi = it->second;
        const int sig = pi.ids[0], internal = pi.ids[1], gnd = pi.ids[2], src = pi.ids[3];
        out->port_zpref[q] = pi.zpref;
        out->port_rows[3 * q + 0] = (int32_t)rowmap[sig];
        out->port_rows[3 * q + 2] = (int32_t)rowmap[gnd];
        if (mode == HVB_MODE_NORTON) {
            out->port_rows[3 * q + 1] = ROW_VIRTUAL_INT;
            if (rowmap[sig] >= 0) add(rowmap[sig], n + q, pi.g, 0);
            if (rowmap[gnd] >= 0) add(rowmap[gnd], n + q, pi.g, 1);
        } else {
            out->port_rows[3 * q + 1] = (int32_t)rowmap[internal];
            if (mode == HVB_MODE_FULL) add(rowmap[src], n + q, 0, 0);
        }
    }
    for (long long r = n_real; r < n; ++r) add(r, r, 0, 0);          // identity padding rows

    // ---- finalise value ids and the two sparse layouts (std::map iterates cells sorted by (row, column))
    const long long nvc = (long long)b.const_vals.size();
    auto fin = [&](long long vid) { return vid >= DYN ? vid - DYN + nvc : vid; };
    std::vector<long long> counts((size_t)(n * ncols), 0);
    out->cell_ent.clear(); out->row_ent.clear();
    out->row_ptr.assign((size_t)n + 1, 0);
    for (auto& kv : cells) {
        const int r = kv.first.first, c = kv.first.second;
        counts[(size_t)r * ncols + c] = (long long)kv.second.size();
        if (c >= (1 << 11)) { *err = "plan too large for the packed row-entry format"; return -1; }
        for (size_t k = 0; k < kv.second.size(); ++k) {
            const long long v = fin(kv.second[k].first);
            const int neg = kv.second[k].second;
            if (v >= (1 << 19)) { *err = "plan too large for the packed row-entry format"; return -1; }
            out->cell_ent.push_back((int32_t)((v << 1) | neg));
            const uint32_t word = (uint32_t)(neg & 1) | ((uint32_t)v << 1) | ((uint32_t)c << 20) | (k + 1 == kv.second.size() ? 0x80000000u : 0u);
            out->row_ent.push_back((int32_t)word);
        }
        out->row_ptr[(size_t)r + 1] += (int32_t)kv.second.size();
    }
    for (long long r = 0; r < n; ++r) out->row_ptr[(size_t)r + 1] += out->row_ptr[(size_t)r];
    out->cell_ptr.assign((size_t)(n * ncols) + 1, 0);
    for (long long i = 0; i < n * ncols; ++i) out->cell_ptr[(size_t)i + 1] = out->cell_ptr[(size_t)i] + (int32_t)counts[(size_t)i];
    for (size_t o = 0; o + HVB_OP_WORDS <= b.ops.size(); o += HVB_OP_WORDS) b.ops[o + 1] += (int32_t)nvc;
    for (size_t o = 0; o + HVB_COOP_WORDS <= b.coops.size(); o += HVB_COOP_WORDS) b.coops[o + 1] += (int32_t)nvc;

    // ---- fast value ops: simple ops whose argument is a real constant or a plain sample column
    out->fop_code.clear(); out->fop_out.clear(); out->fop_arg.clear(); out->ops.clear();
    for (size_t o = 0; o + HVB_OP_WORDS <= b.ops.size(); o += HVB_OP_WORDS) {
        const int code = b.ops[o], outv = b.ops[o + 1], pi = b.ops[o + 2];
        const bool simple = code == OP_COPY || code == OP_RECIP || code == OP_CAP || code == OP_IND;
        const int kind = simple ? b.prefs[2 * (size_t)pi] : -1, idx = simple ? b.prefs[2 * (size_t)pi + 1] : 0;
        if (kind == PK_CONST && b.consts[idx].im == 0.0) { out->fop_code.push_back(code); out->fop_out.push_back(outv - (int32_t)nvc); out->fop_arg.push_back(b.consts[idx].re); }
        else if (kind == PK_SAMPLE) { out->fop_code.push_back(code | 4); out->fop_out.push_back(outv - (int32_t)nvc); out->fop_arg.push_back((double)idx); }
        else out->ops.insert(out->ops.end(), b.ops.begin() + o, b.ops.begin() + o + HVB_OP_WORDS);
    }
    // ---- closed-form epilogue
    out->epi_simple = 0;
    out->epi_port_of_row.assign((size_t)std::max<long long>(n, 1), -1);
    out->epi_scale.assign((size_t)std::max(P, 1) * std::max(P, 1), 1.0);
    if (mode == HVB_MODE_NORTON && P > 0) {
        std::vector<double> z;
        bool ok = true;
        for (int q = 0; q < P && ok; ++q) {
            const int kind = b.prefs[2 * (size_t)out->port_zpref[q]], idx = b.prefs[2 * (size_t)out->port_zpref[q] + 1];
            if (kind != PK_CONST || b.consts[idx].im != 0.0 || !(b.consts[idx].re > 0.0)) { ok = false; break; }
            z.push_back(b.consts[idx].re);
            const int sig = out->port_rows[3 * q], gnd = out->port_rows[3 * q + 2];
            if (gnd != ROW_GROUND || sig < 0 || out->epi_port_of_row[sig] != -1) { ok = false; break; }
            out->epi_port_of_row[sig] = q;
        }
        if (ok) {
            out->epi_simple = 1;
            for (int j = 0; j < P; ++j) for (int i = 0; i < P; ++i) out->epi_scale[(size_t)j * P + i] = std::sqrt(std::fabs(z[i])) / std::sqrt(std::fabs(z[j]));
        } else std::fill(out->epi_port_of_row.begin(), out->epi_port_of_row.end(), -1);
    }
    out->mode = mode; out->n = (int)n; out->ncols = (int)ncols; out->P = P; out->n_real = (int)n_real; out->band = band;
    out->n_dyn = b.dyn_count; out->coop_scratch = b.coop_scratch; out->max_nport = b.max_nport;
    out->n_tables = b.n_tables; out->n_sample_cols = b.n_sample_cols;
    out->consts.resize(b.consts.size() * 2);
    for (size_t i = 0; i < b.consts.size(); ++i) { out->consts[2 * i] = b.consts[i].re; out->consts[2 * i + 1] = b.consts[i].im; }
    out->prefs = b.prefs;
    out->const_vals.resize(b.const_vals.size() * 2);
    for (size_t i = 0; i < b.const_vals.size(); ++i) { out->const_vals[2 * i] = b.const_vals[i].re; out->const_vals[2 * i + 1] = b.const_vals[i].im; }
    out->coops = b.coops;
    out->table_pref = b.table_pref; out->table_slot = b.table_slot;
    return 0;
}

void StampPlan::fill_desc(const hvb_stamp_table& t, hvb_plan_desc* d, int kernel_hint) const {
    memset(d, 0, sizeof(*d));
    d->n = n; d->ncols = ncols; d->P = P;
    d->n_consts = (int32_t)consts.size() / 2; d->n_prefs = (int32_t)prefs.size() / 2;
    d->n_const_vals = (int32_t)const_vals.size() / 2; d->n_dyn = n_dyn;
    d->n_ops = (int32_t)(ops.size() / HVB_OP_WORDS); d->n_coops = (int32_t)(coops.size() / HVB_COOP_WORDS);
    d->n_samples_cols = n_sample_cols; d->coop_scratch = coop_scratch;
    d->n_traces = t.n_traces; d->n_knots = t.n_knots; d->n_coefs = t.n_coefs;
    d->n_row_ent = (int32_t)row_ent.size(); d->n_fast_ops = (int32_t)fop_code.size();
    d->epi_simple = epi_simple; d->kernel_hint = kernel_hint;
    static const double zero2[2] = {0.0, 0.0};
    static const int32_t zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    d->const_vals = const_vals.data();
    d->ops = ops.empty() ? zero8 : ops.data();
    d->coops = coops.empty() ? zero8 : coops.data();
    d->cell_ptr = cell_ptr.data(); d->cell_ent = cell_ent.empty() ? zero8 : cell_ent.data();
    d->row_ptr = row_ptr.data(); d->row_ent = row_ent.empty() ? zero8 : row_ent.data();
    d->fop_code = fop_code.empty() ? zero8 : fop_code.data(); d->fop_out = fop_out.empty() ? zero8 : fop_out.data();
    d->fop_arg = fop_arg.empty() ? zero2 : fop_arg.data();
    d->epi_port_of_row = epi_port_of_row.data(); d->epi_scale = epi_scale.data();
    d->port_rows = port_rows.empty() ? zero8 : port_rows.data(); d->port_zpref = port_zpref.empty() ? zero8 : port_zpref.data();
    d->spl_t = t.spl_t ? t.spl_t : zero2; d->spl_c = t.spl_c ? t.spl_c : zero2;
    d->spl_trace = t.spl_trace ? t.spl_trace : zero8; d->spl_fill = t.spl_fill ? t.spl_fill : zero2; d->spl_range = t.spl_range ? t.spl_range : zero2;
}
