// hvb_band.cu -- kernel C: one THREAD per (sample, frequency) point, banded LU with partial pivoting.
//
// Large netlists are chains: after the plan compiler's reverse Cuthill-McKee ordering the MNA matrix
// of a ladder / transmission-line network has a half bandwidth B of 1..8 while n is in the hundreds,
// so a dense factorisation spends > 99 % of its flops on structural zeros.  Row-oriented banded
// elimination with partial pivoting (the LAPACK zgbtrf recurrence: U grows to bandwidth 2B, pivots
// are searched in the B rows below the diagonal -- the same pivots a dense solve picks, because the
// rest of the column is exactly zero) needs n*B*(2B+P) complex FMAs instead of n^3/3.  The
// recurrence is sequential in k, so the parallel axis is the batch: every thread owns one point.
//
// Workspace (global memory, element e of thread t at W[e*T + t], T = threads of the launch, so a
// warp's access to "its" element e is one 512-byte coalesced line):
//     band rows  [n + 2B][LD],  LD = 3B+1 + PP:  A[r][c] at column c - r + B, RHS i at column 3B+1+i
//     values     [n_dyn]
// The kernel is bound by that traffic (HBM / L2), not by FP64.
// Same results as solvers/spfn.py:140-176 for the same netlist (different pivot order than the dense
// kernels only where the row order differs).
#include "hvb_device.cuh"
#include "hvb_kernels.h"

#define HVB_BAND_THREADS 128
#ifndef HVB_BANDF_MINB1
#define HVB_BANDF_MINB1 5
#endif

struct BandView {
    cplx* __restrict__ base;      // W + t
    long long T;                  // element stride
    int LD;
    __device__ __forceinline__ cplx& at(int r, int c) const { return base[((long long)r * LD + c) * T]; }
};

__device__ __forceinline__ cplx band_port_volt(const HvbDev& D, const BandView& W, int xcol, int q, int which, int i) {
    const int r = __ldg(D.port_rows + 3 * q + which);
    if (r >= 0) return W.at(r, xcol + i);
    if (r == ROW_GROUND) return cmk(0.0, 0.0);
    const int rg = __ldg(D.port_rows + 3 * q + 2);
    cplx v = rg >= 0 ? W.at(rg, xcol + i) : cmk(0.0, 0.0);
    if (q == i) v.x += 1.0;
    return v;
}

template <int B>
__global__ void __launch_bounds__(HVB_BAND_THREADS) hvb_band_kernel(const HvbDev D) {
    constexpr int LDA = 3 * B + 1;
    const int n = D.n, PP = D.PP, P = D.P;
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    BandView W{D.W + t, T, LDA + PP};
    cplx* __restrict__ Vt = D.W + (long long)(n + 2 * B) * W.LD * T + t;   // values: Vt[v*T]
    const long long total = D.nS * D.nF;
    unsigned bad = 0;

    for (long long q = t; q < total; q += T) {
        const long long s = q / D.nF, fi = q - s * D.nF;
        const double f = __ldg(D.f + fi);
        if (D.f32 && s == 0) D.f32[fi] = (float)f;
        const double* __restrict__ srow = D.samples + s * D.nR;

        // ---- value program ----------------------------------------------------------------------
        {
            const double w = __dmul_rn(HVB_TWOPI, f);
            for (int o = 0; o < D.nfast; ++o) {            // same arithmetic as eval_fast_ops
                const int code = __ldg(D.fop_code + o);
                double p = __ldg(D.fop_arg + o);
                if (code & 4) p = srow[(int)p];
                const double tt = (code & 2) ? __dmul_rn(w, p) : p;
                const double v = (code & 1) ? 1.0 / tt : tt;
                const bool imag = (code & 2) != 0;
                Vt[(long long)__ldg(D.fop_out + o) * T] = cmk(imag ? 0.0 : v, imag ? ((code & 1) ? -v : v) : 0.0);
            }
            for (int o = 0; o < D.nops; ++o) {
                const int* __restrict__ op = D.ops + o * HVB_OP_WORDS;
                const int out = __ldg(op + 1) - D.nvc;
                cplx tmp[9];
                eval_simple_op(D, op, f, fi, srow, tmp - out);     // writes tmp[0] (tmp[0..8] for a line)
                const int cnt = __ldg(op) == OP_TLINE ? 9 : 1;
                for (int e = 0; e < cnt; ++e) Vt[(long long)(out + e) * T] = tmp[e];
            }
        }

        // ---- assembly: every band cell written exactly once (row entries are sorted by column) ----
        for (int r = 0; r < n; ++r) {
            int e = __ldg(D.row_ptr + r);
            const int hi = __ldg(D.row_ptr + r + 1);
            int ecol = e < hi ? (int)(((unsigned)__ldg(D.row_ent + e) >> 20) & 0x7FFu) : -1;
            for (int bc = 0; bc < W.LD; ++bc) {
                int col = bc < LDA ? r - B + bc : n + bc - LDA;
                if (bc < LDA && col >= n) col = -2;            // band slot beyond the last column, not an RHS column
                cplx acc = cmk(0.0, 0.0);
                while (e < hi && ecol == col) {                // e < hi: `col` can itself be -1 left of the band's first column
                    const int word = __ldg(D.row_ent + e);
                    const int vid = (word >> 1) & 0x7FFFF;
                    const cplx v = vid < D.nvc ? D.const_vals[vid] : Vt[(long long)(vid - D.nvc) * T];
                    if (word & 1) { acc.x -= v.x; acc.y -= v.y; } else { acc.x += v.x; acc.y += v.y; }
                    ++e;
                    ecol = e < hi ? (int)(((unsigned)__ldg(D.row_ent + e) >> 20) & 0x7FFu) : -1;
                }
                W.at(r, bc) = acc;
            }
        }
        for (int r = n; r < n + 2 * B; ++r)
            for (int bc = 0; bc < W.LD; ++bc) W.at(r, bc) = cmk(0.0, 0.0);

        // ---- banded elimination with partial pivoting ----------------------------------------------
        double pmin = 1e300, pmax = 0.0;
        for (int k = 0; k < n; ++k) {
            int p = 0;
            double best = -1.0;
#pragma unroll
            for (int i = 0; i <= B; ++i) {
                const cplx c = W.at(k + i, B - i);
                const double m = fabs(c.x) + fabs(c.y);
                if (m > best) { best = m; p = i; }          // first largest, like the dense kernels
            }
            pmin = fmin(pmin, best);
            pmax = fmax(pmax, best);
            if (p != 0) {                                   // physical row exchange inside the band
#pragma unroll
                for (int j = 0; j <= 2 * B; ++j) {
                    const cplx a = W.at(k, B + j), b = W.at(k + p, B + j - p);
                    W.at(k, B + j) = b;
                    W.at(k + p, B + j - p) = a;
                }
                for (int r = 0; r < PP; ++r) {
                    const cplx a = W.at(k, LDA + r), b = W.at(k + p, LDA + r);
                    W.at(k, LDA + r) = b;
                    W.at(k + p, LDA + r) = a;
                }
            }
            const cplx inv = crecip_fast(W.at(k, B));
            W.at(k, B) = inv;                               // the back substitution multiplies by it
            cplx u[2 * B + 1];
#pragma unroll
            for (int j = 1; j <= 2 * B; ++j) u[j] = W.at(k, B + j);
#pragma unroll
            for (int i = 1; i <= B; ++i) {
                const cplx a = W.at(k + i, B - i);
                if (a.x != 0.0 || a.y != 0.0) {
                    const cplx l = cmul(a, inv);
#pragma unroll
                    for (int j = 1; j <= 2 * B; ++j) {
                        cplx v = W.at(k + i, B - i + j);
                        cfnma(v, l, u[j]);
                        W.at(k + i, B - i + j) = v;
                    }
                    for (int r = 0; r < PP; ++r) {
                        cplx v = W.at(k + i, LDA + r);
                        cfnma(v, l, W.at(k, LDA + r));
                        W.at(k + i, LDA + r) = v;
                    }
                }
            }
        }
        if (!(pmin >= 2.2250738585072014e-308)) bad |= HVB_STATUS_ZERO_PIVOT_BIT;    // also a NaN column (best stays -1)
        if (!(pmax <= 1.7976931348623157e308)) bad |= HVB_STATUS_NONFINITE_BIT;

        // ---- back substitution (rows n .. n+2B-1 of the RHS block are zero) ------------------------
        for (int k = n - 1; k >= 0; --k) {
            const cplx inv = W.at(k, B);
            cplx u[2 * B + 1];
#pragma unroll
            for (int j = 1; j <= 2 * B; ++j) u[j] = W.at(k, B + j);
            for (int r = 0; r < PP; ++r) {
                cplx acc = W.at(k, LDA + r);
#pragma unroll
                for (int j = 1; j <= 2 * B; ++j) cfnma(acc, u[j], W.at(k + j, LDA + r));
                W.at(k, LDA + r) = cmul(acc, inv);
            }
        }

        // ---- S epilogue (solvers/spfn.py:159-176) ---------------------------------------------------
        for (int i = 0; i < P; ++i) {
            const cplx Zi = eval_param(D.pv, __ldg(D.port_zpref + i), f, fi, srow);
            const cplx Vi1 = band_port_volt(D, W, LDA, i, 0, i), Vi3 = band_port_volt(D, W, LDA, i, 1, i),
                       Vi2 = band_port_volt(D, W, LDA, i, 2, i);
            const cplx den = csub(csub(Vi1, Vi2), cdiv_plain(cmul(Zi, csub(Vi1, Vi3)), Zi));
            for (int j = 0; j < P; ++j) {
                const cplx Zj = eval_param(D.pv, __ldg(D.port_zpref + j), f, fi, srow);
                const cplx Vo1 = band_port_volt(D, W, LDA, j, 0, i), Vo3 = band_port_volt(D, W, LDA, j, 1, i),
                           Vo2 = band_port_volt(D, W, LDA, j, 2, i);
                const cplx num = cadd(csub(Vo1, Vo2), cdiv_plain(cmul(cconj(Zj), csub(Vo1, Vo3)), Zj));
                cplx sv = cdiv_plain(num, den);
                if (Zi.x != Zj.x) sv = cscale(sv, sqrt(fabs(Zi.x)) / sqrt(fabs(Zj.x)));
                if (!(isfinite(sv.x) && isfinite(sv.y))) bad |= HVB_STATUS_NONFINITE_BIT;
                D.S[((s * P + j) * P + i) * D.ldS + fi] = sv;
            }
        }
    }
    if (bad && D.status) atomicOr(D.status, (int)bad);
}

// ------------------------------------------------------------------------------------------------
// Fused variant (B <= 4): nothing but finished pivot rows ever reaches global memory.
//   * the matrix part of the B+1 rows live at elimination step k (columns k .. k+2B) stays in
//     REGISTERS; their right-hand sides sit in a shared-memory ring (row exchanges and the slide of
//     the window only permute ring indices);
//   * row k+B is assembled straight into the window when step k starts, from per-point values that
//     are evaluated just in time into a small pool of shared-memory slots (see BandVals);
//   * every finished pivot row is appended once to a per-thread record stream
//     [inv(u_kk), u_k,k+1 .. u_k,k+2B, y_k,0 .. y_k,act(k)-1] -- act(k) = right-hand-side columns that
//     can be non-zero in rows <= k+B, a prefix because a port's excitation enters at its own row --
//     and read once, back to front, by the substitution, whose last 2B solution rows live in the
//     same shared-memory ring; only the rows the S epilogue needs (port nodes) are kept.
// HBM traffic per point: one write + one read of n*(2B+1+<act>) complex words.
// ------------------------------------------------------------------------------------------------
struct BandVals {
    cplx* sm;
    cplx* spill;
    long long T;
    int nsm;
    __device__ __forceinline__ cplx load(int slot) const {
        return slot < nsm ? sm[slot * HVB_BAND_THREADS] : spill[(long long)(slot - nsm) * T];
    }
    __device__ __forceinline__ void store(int slot, cplx v) const {
        if (slot < nsm) sm[slot * HVB_BAND_THREADS] = v; else spill[(long long)(slot - nsm) * T] = v;
    }
};

// Two-port S or Y block evaluated by ONE thread (base/nport.py:28-39; the lane-cooperative version is
// coop_nport in hvb_warp.cuh): S -> Y = (1/Z0)(I - S)(I + S)^-1 as a solve of (I+S)^T X^T = (I-S)^T with
// partial pivoting, then the indefinite 3x3 block with the reference node's row / column (-row sums,
// -column sums, +total).  o[r*3 + c]; returns status bits.
// (out of line and fed a ParamView copy: only netlists with such blocks pay for it)
static __device__ __noinline__ unsigned band_eval_twoport(const ParamView pv, const int* __restrict__ op, double f, long long fi,
                                                          const double* __restrict__ srow, cplx* __restrict__ o) {
    const int code = __ldg(op), first = __ldg(op + 3);
    unsigned bad = 0;
    cplx y[2][2];                                           // y[i][j]
    if (code == COOP_NPORT_Y) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) y[i][j] = eval_param(pv, first + i * 2 + j, f, fi, srow);
    } else {
        cplx w[2][4];                                       // rows of [(I+S)^T | (I-S)^T]
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const cplx s = eval_param(pv, first + i * 2 + j, f, fi, srow);
                const double d = (i == j) ? 1.0 : 0.0;
                w[j][i] = cmk(d + s.x, s.y);
                w[j][2 + i] = cmk(d - s.x, -s.y);
            }
        // column 0: pivot row p, the other row q
        const bool swap = fabs(w[1][0].x) + fabs(w[1][0].y) > fabs(w[0][0].x) + fabs(w[0][0].y);
        if (swap) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { const cplx t = w[0][c]; w[0][c] = w[1][c]; w[1][c] = t; }
        }
        const double m0 = fabs(w[0][0].x) + fabs(w[0][0].y);
        const cplx inv0 = crecip_fast(w[0][0]);
        const cplx l = cmul(w[1][0], inv0);
#pragma unroll
        for (int c = 1; c < 4; ++c) cfnma(w[1][c], l, w[0][c]);
        const double m1 = fabs(w[1][1].x) + fabs(w[1][1].y);
        if (!(m0 >= 2.2250738585072014e-308) || !(m1 >= 2.2250738585072014e-308)) bad |= HVB_STATUS_ZERO_PIVOT_BIT;
        const cplx inv1 = crecip_fast(w[1][1]);
        // back substitution: X^T rows, X^T[k][c] = X[c][k]
        cplx xt[2][2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            xt[1][c] = cmul(w[1][2 + c], inv1);
            cplx t = w[0][2 + c];
            cfnma(t, w[0][1], xt[1][c]);
            xt[0][c] = cmul(t, inv0);
        }
        const cplx iz = crecip_np(pv.consts[__ldg(op + 4)]);     // 1/Z0 (nport.py:32)
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int c = 0; c < 2; ++c) y[c][k] = (iz.y == 0.0) ? cscale(xt[k][c], iz.x) : cmul(iz, xt[k][c]);
    }
    cplx tot = cmk(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const cplx rs = cadd(y[i][0], y[i][1]), cs = cadd(y[0][i], y[1][i]);
        o[i * 3 + 0] = y[i][0]; o[i * 3 + 1] = y[i][1];
        o[i * 3 + 2] = cneg(rs);
        o[2 * 3 + i] = cneg(cs);
        tot = cadd(tot, rs);
    }
    o[8] = tot;
    return bad;
}

// evaluate the value ops scheduled before row r (hvb_api.cu: build_band_schedule)
template <bool TWOPORT>
__device__ __forceinline__ unsigned band_run_schedule(const HvbDev& D, int r, double f, long long fi,
                                                      const double* __restrict__ srow, const BandVals& vals) {
    unsigned bad = 0;
    const int lo = __ldg(D.band_sched_ptr + r), hi = __ldg(D.band_sched_ptr + r + 1);
    const double w = __dmul_rn(HVB_TWOPI, f);
    for (int i = lo; i < hi; ++i) {
        const int ref = __ldg(D.band_sched_op + i);
        if (ref >= 0) {                                     // fast op (same arithmetic as eval_fast_ops)
            const int code = __ldg(D.fop_code + ref);
            double p = __ldg(D.fop_arg + ref);
            if (code & 4) p = srow[(int)p];
            const double tt = (code & 2) ? __dmul_rn(w, p) : p;
            const double v = (code & 1) ? 1.0 / tt : tt;
            const bool imag = (code & 2) != 0;
            vals.store(__ldg(D.band_fslot + ref), cmk(imag ? 0.0 : v, imag ? ((code & 1) ? -v : v) : 0.0));
        } else {
            const int o = ~ref;                             // simple op, or nops + index of a two-port block
            cplx tmp[9];
            if (TWOPORT && o >= D.nops) {
                bad |= band_eval_twoport(D.pv, D.coops + (o - D.nops) * HVB_COOP_WORDS, f, fi, srow, tmp);
#pragma unroll
                for (int e = 0; e < 9; ++e) {
                    const int sl = __ldg(D.band_oslot + o * 9 + e);
                    if (sl >= 0) vals.store(sl, tmp[e]);
                }
                continue;
            }
            const int* __restrict__ op = D.ops + o * HVB_OP_WORDS;
            const int out = __ldg(op + 1) - D.nvc;
            const bool line = __ldg(op) == OP_TLINE;
            if (line) {
                const cplx Z0 = eval_param(D.pv, __ldg(op + 2), f, fi, srow);
                const cplx beta = eval_param(D.pv, __ldg(op + 3), f, fi, srow);
                eval_tline_fast(Z0, beta, eval_param(D.pv, __ldg(op + 4), f, fi, srow).x, tmp);
            } else {
                eval_simple_op(D, op, f, fi, srow, tmp - out);
            }
            const int cnt = line ? 9 : 1;
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                if (e < cnt) {
                    const int sl = __ldg(D.band_oslot + o * 9 + e);
                    if (sl >= 0) vals.store(sl, tmp[e]);
                }
            }
        }
    }
    return bad;
}

// Assemble system row r: matrix cells (columns k0 .. k0+2B) into registers a[], right-hand-side cells
// into the shared-memory ring row yr[] (element i at yr[i*THREADS]).  D.band_ent is row_ent, sorted
// by column, with per-point value ids replaced by nvc + slot.
template <int B>
__device__ __forceinline__ void band_assemble_row(const HvbDev& D, int r, int k0, const BandVals& vals,
                                                  cplx (&a)[2 * B + 1], cplx* __restrict__ yr) {
    int e = __ldg(D.row_ptr + r);
    const int hi = __ldg(D.row_ptr + r + 1);
    int word = e < hi ? __ldg(D.band_ent + e) : 0;
    int ecol = e < hi ? (int)(((unsigned)word >> 20) & 0x7FFu) : -1;
    auto value = [&]() {
        const int vid = (word >> 1) & 0x7FFFF;
        cplx v = vid < D.nvc ? D.const_vals[vid] : vals.load(vid - D.nvc);
        if (word & 1) v = cneg(v);
        ++e;
        word = e < hi ? __ldg(D.band_ent + e) : 0;
        ecol = e < hi ? (int)(((unsigned)word >> 20) & 0x7FFu) : -1;
        return v;
    };
#pragma unroll
    for (int c = 0; c <= 2 * B; ++c) {
        const int col = k0 + c < D.n ? k0 + c : -2;
        cplx acc = cmk(0.0, 0.0);
        while (ecol == col) acc = cadd(acc, value());
        a[c] = acc;
    }
    for (int i = 0; i < D.P; ++i) yr[i * HVB_BAND_THREADS] = cmk(0.0, 0.0);   // all of it: act(k) grows with k
    while (e < hi) {                                        // what is left are right-hand-side cells (col >= n)
        const int i = ecol - D.n;
        const cplx v = value();
        yr[i * HVB_BAND_THREADS] = cadd(yr[i * HVB_BAND_THREADS], v);
    }
}

// TWOPORT: the plan holds two-port S / Y blocks (evaluated per thread); kept out of the common instantiation
// because the extra call path costs the hot loop registers (c5: 17 -> 23 ms when it was unconditional)
template <int B, bool TWOPORT>
__global__ void __launch_bounds__(HVB_BAND_THREADS, B == 1 ? HVB_BANDF_MINB1 : B == 2 ? 4 : B == 3 ? 3 : 2) hvb_bandf_kernel(const HvbDev D) {
    constexpr int WD = 2 * B + 1, NR = B + 1, NX = 2 * B, NRING = NR > NX ? NR : NX;
    const int n = D.n, P = D.P;
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    cplx* __restrict__ Rt = D.W + t;                                   // record stream, element e at Rt[e*T]
    cplx* __restrict__ Xs = Rt + (long long)D.band_rec_words * T;     // x(slot, r) of the rows the epilogue reads
    extern __shared__ __align__(16) unsigned char hvb_band_smem[];
    cplx* __restrict__ ring = reinterpret_cast<cplx*>(hvb_band_smem) + threadIdx.x;   // ring row i, column r: ring[(i*P + r)*THREADS]
    const BandVals vals{ring + NRING * P * HVB_BAND_THREADS, Xs + (long long)D.band_nslots * P * T, T, D.band_nsm};
    const int ring_row = P * HVB_BAND_THREADS;
    const long long total = D.nS * D.nF;
    unsigned bad = 0;

    for (long long q = t; q < total; q += T) {
        const long long s = q / D.nF, fi = q - s * D.nF;
        const double f = __ldg(D.f + fi);
        if (D.f32 && s == 0) D.f32[fi] = (float)f;
        const double* __restrict__ srow = D.samples + s * D.nR;

        // ---- elimination ----------------------------------------------------------------------------
        cplx A[NR][WD];
        int yrow[NR];                                       // ring row holding the right-hand sides of window row i
#pragma unroll
        for (int i = 0; i < NR; ++i) yrow[i] = i;
        {
#pragma unroll
            for (int i = 0; i < B; ++i) {
                if (i < n) {
                    bad |= band_run_schedule<TWOPORT>(D, i, f, fi, srow, vals);
                    band_assemble_row<B>(D, i, 0, vals, A[i], ring + yrow[i] * ring_row);
                } else {
#pragma unroll
                    for (int c = 0; c < WD; ++c) A[i][c] = cmk(0.0, 0.0);
                    for (int r = 0; r < P; ++r) ring[yrow[i] * ring_row + r * HVB_BAND_THREADS] = cmk(0.0, 0.0);
                }
            }
        }
        double pmin = 1e300, pmax = 0.0;
        cplx* __restrict__ wp = Rt;
        for (int k = 0; k < n; ++k) {
            const int na = __ldg(D.band_act + k);
            cplx* __restrict__ ynew = ring + yrow[B] * ring_row;
            if (k + B < n) {
                bad |= band_run_schedule<TWOPORT>(D, k + B, f, fi, srow, vals);
                band_assemble_row<B>(D, k + B, k, vals, A[B], ynew);
            } else {
#pragma unroll
                for (int c = 0; c < WD; ++c) A[B][c] = cmk(0.0, 0.0);
                for (int r = 0; r < P; ++r) ynew[r * HVB_BAND_THREADS] = cmk(0.0, 0.0);
            }
            int p = 0;
            double best = fabs(A[0][0].x) + fabs(A[0][0].y);
            if (!(best >= 0.0)) best = -1.0;                 // NaN never wins
#pragma unroll
            for (int i = 1; i <= B; ++i) {
                const double m = fabs(A[i][0].x) + fabs(A[i][0].y);
                if (m > best) { best = m; p = i; }
            }
            pmin = fmin(pmin, best);
            pmax = fmax(pmax, best);
#pragma unroll
            for (int i = 1; i <= B; ++i) {
                if (p == i) {
#pragma unroll
                    for (int c = 0; c < WD; ++c) { const cplx x = A[0][c]; A[0][c] = A[i][c]; A[i][c] = x; }
                    const int x = yrow[0]; yrow[0] = yrow[i]; yrow[i] = x;
                }
            }
            const cplx inv = crecip_fast(A[0][0]);
            const cplx* __restrict__ y0 = ring + yrow[0] * ring_row;
            *wp = inv; wp += T;
#pragma unroll
            for (int c = 1; c < WD; ++c) { *wp = A[0][c]; wp += T; }
            for (int r = 0; r < na; ++r) { *wp = y0[r * HVB_BAND_THREADS]; wp += T; }
#pragma unroll
            for (int i = 1; i <= B; ++i) {
                const cplx l = cmul(A[i][0], inv);
#pragma unroll
                for (int c = 1; c < WD; ++c) cfnma(A[i][c], l, A[0][c]);
                cplx* __restrict__ yi = ring + yrow[i] * ring_row;
#pragma unroll 4
                for (int r = 0; r < na; ++r) {
                    cplx v = yi[r * HVB_BAND_THREADS];
                    cfnma(v, l, y0[r * HVB_BAND_THREADS]);
                    yi[r * HVB_BAND_THREADS] = v;
                }
            }
            // slide: row i becomes row i-1, column c becomes column c-1; the pivot row's ring row is recycled
            const int freed = yrow[0];
#pragma unroll
            for (int i = 1; i <= B; ++i) {
#pragma unroll
                for (int c = 1; c < WD; ++c) A[i - 1][c - 1] = A[i][c];
                A[i - 1][WD - 1] = cmk(0.0, 0.0);
                yrow[i - 1] = yrow[i];
            }
            yrow[B] = freed;
        }
        if (!(pmin >= 2.2250738585072014e-308)) bad |= HVB_STATUS_ZERO_PIVOT_BIT;
        if (!(pmax <= 1.7976931348623157e308)) bad |= HVB_STATUS_NONFINITE_BIT;

        // ---- back substitution: x_{k+1} .. x_{k+2B} in the shared-memory ring ------------------------
        {
            for (int i = 0; i < NX * P; ++i) ring[i * HVB_BAND_THREADS] = cmk(0.0, 0.0);
            int xrow[NX + 1];                               // xrow[j] = ring row of x_{k+j}; xrow[0] is the row x_k goes to
#pragma unroll
            for (int j = 0; j <= NX; ++j) xrow[j] = j % NX;
            const cplx* __restrict__ rp = wp;               // end of the stream
            const cplx* __restrict__ pp = wp;               // prefetch cursor, band_pfb records ahead
            int kp = n - 1;
            for (int d = 0; d < D.band_pfb && kp >= D.band_min_row; ++d, --kp) {
                const int rs = WD + __ldg(D.band_act + kp);
                pp -= rs * T;
                const cplx* __restrict__ pe = pp;
                for (int e = 0; e < rs; ++e, pe += T) asm volatile("prefetch.global.L2 [%0];" ::"l"(pe));
            }
            for (int k = n - 1; k >= D.band_min_row; --k) {
                if (D.band_pfb > 0 && kp >= D.band_min_row) {
                    const int rs = WD + __ldg(D.band_act + kp);
                    pp -= rs * T;
                    const cplx* __restrict__ pe = pp;
                    for (int e = 0; e < rs; ++e, pe += T) asm volatile("prefetch.global.L2 [%0];" ::"l"(pe));
                    --kp;
                }
                const int na = __ldg(D.band_act + k);
                rp -= (WD + na) * T;
                cplx u[WD];
                const cplx* __restrict__ yp = rp;
#pragma unroll
                for (int c = 0; c < WD; ++c, yp += T) u[c] = *yp;
                const int slot = __ldg(D.band_slot_of_row + k);
                cplx* __restrict__ xo = ring + xrow[0] * ring_row;
                const cplx* xj[NX + 1];
#pragma unroll
                for (int j = 1; j <= NX; ++j) xj[j] = ring + xrow[j] * ring_row;
                cplx* __restrict__ xs = Xs + (long long)(slot < 0 ? 0 : slot) * P * T;
                // four right-hand sides at a time: their record loads are issued together
                for (int r0 = 0; r0 < P; r0 += 4) {
                    cplx acc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i] = (r0 + i < na) ? *yp : cmk(0.0, 0.0);
                        if (r0 + i < na) yp += T;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (r0 + i < P) {
                            const int off = (r0 + i) * HVB_BAND_THREADS;
#pragma unroll
                            for (int j = 1; j <= NX; ++j) cfnma(acc[i], u[j], xj[j][off]);
                            acc[i] = cmul(acc[i], u[0]);
                            xo[off] = acc[i];                // xrow[0] == xrow[NX]: x_{k+2B} was read just above
                            if (slot >= 0) { *xs = acc[i]; xs += T; }
                        }
                    }
                }
                // x_k becomes x_{(k-1)+1}
#pragma unroll
                for (int j = NX; j >= 1; --j) xrow[j] = xrow[j - 1];
                xrow[0] = xrow[NX];
            }
        }

        // ---- S epilogue (solvers/spfn.py:159-176) ---------------------------------------------------
        auto volt = [&](int qp, int which, int i) {
            const int r = __ldg(D.port_rows + 3 * qp + which);
            if (r >= 0) return Xs[((long long)__ldg(D.band_slot_of_row + r) * P + i) * T];
            if (r == ROW_GROUND) return cmk(0.0, 0.0);
            const int rg = __ldg(D.port_rows + 3 * qp + 2);
            cplx v = rg >= 0 ? Xs[((long long)__ldg(D.band_slot_of_row + rg) * P + i) * T] : cmk(0.0, 0.0);
            if (qp == i) v.x += 1.0;
            return v;
        };
        for (int i = 0; i < P; ++i) {
            const cplx Zi = eval_param(D.pv, __ldg(D.port_zpref + i), f, fi, srow);
            const cplx Vi1 = volt(i, 0, i), Vi3 = volt(i, 1, i), Vi2 = volt(i, 2, i);
            const cplx den = csub(csub(Vi1, Vi2), cdiv_plain(cmul(Zi, csub(Vi1, Vi3)), Zi));
            for (int j = 0; j < P; ++j) {
                const cplx Zj = eval_param(D.pv, __ldg(D.port_zpref + j), f, fi, srow);
                const cplx Vo1 = volt(j, 0, i), Vo3 = volt(j, 1, i), Vo2 = volt(j, 2, i);
                const cplx num = cadd(csub(Vo1, Vo2), cdiv_plain(cmul(cconj(Zj), csub(Vo1, Vo3)), Zj));
                cplx sv = cdiv_plain(num, den);
                if (Zi.x != Zj.x) sv = cscale(sv, sqrt(fabs(Zi.x)) / sqrt(fabs(Zj.x)));
                if (!(isfinite(sv.x) && isfinite(sv.y))) bad |= HVB_STATUS_NONFINITE_BIT;
                D.S[((s * P + j) * P + i) * D.ldS + fi] = sv;
            }
        }
    }
    if (bad && D.status) atomicOr(D.status, (int)bad);
}

const void* hvb_bandf_kernel_ptr(int b, int twoport) {
    switch (b) {
        case 1: return twoport ? (const void*)hvb_bandf_kernel<1, true> : (const void*)hvb_bandf_kernel<1, false>;
        case 2: return twoport ? (const void*)hvb_bandf_kernel<2, true> : (const void*)hvb_bandf_kernel<2, false>;
        case 3: return twoport ? (const void*)hvb_bandf_kernel<3, true> : (const void*)hvb_bandf_kernel<3, false>;
        case 4: return twoport ? (const void*)hvb_bandf_kernel<4, true> : (const void*)hvb_bandf_kernel<4, false>;
    }
    return nullptr;
}
// shared-memory complex words per thread: the right-hand-side / solution ring (value slots come on top)
int hvb_bandf_ring_words(int b, int p) { return (b + 1 > 2 * b ? b + 1 : 2 * b) * p; }

const void* hvb_band_kernel_ptr(int b) {
    switch (b) {
        case 1: return (const void*)hvb_band_kernel<1>;
        case 2: return (const void*)hvb_band_kernel<2>;
        case 3: return (const void*)hvb_band_kernel<3>;
        case 4: return (const void*)hvb_band_kernel<4>;
        case 6: return (const void*)hvb_band_kernel<6>;
        case 8: return (const void*)hvb_band_kernel<8>;
    }
    return nullptr;
}
int hvb_band_threads() { return HVB_BAND_THREADS; }
int hvb_band_round_up(int b) { return b <= 4 ? (b < 1 ? 1 : b) : b <= 6 ? 6 : b <= 8 ? 8 : 0; }
size_t hvb_band_words_per_thread(int b, int n, int pp, int n_dyn) {
    return (size_t)(n + 2 * b) * (size_t)(3 * b + 1 + pp) + (size_t)(n_dyn > 0 ? n_dyn : 1);
}
