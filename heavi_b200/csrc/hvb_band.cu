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
#ifndef HVB_BAND_PREFETCH
#define HVB_BAND_PREFETCH 2            // rows of values prefetched ahead of the elimination front
#endif
#ifndef HVB_BAND_PREFETCH_BACK
#define HVB_BAND_PREFETCH_BACK 6       // pivot-row records prefetched (to L2) ahead of the back substitution
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
                while (ecol == col) {
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
// Fused variant for B <= 2: the B+1 rows that are live at elimination step k (columns k .. k+2B and
// the right-hand sides) stay in REGISTERS.  Row k+B is assembled straight into the window when step
// k starts, every finished pivot row is written to global memory exactly once ([inv(u_kk), u_k,k+1..,
// y_k]), and the back substitution reads it exactly once while the last 2B solution rows slide
// through registers; only the rows the S epilogue needs (port nodes) are written back.  Traffic per
// point drops from ~4 passes over the band to one write + one read.
// ------------------------------------------------------------------------------------------------
template <int B, int PP>
struct BandRow {
    cplx a[2 * B + 1];
    cplx y[PP];
};

// pull the per-point values row r sums into L1 ahead of its assembly: the walker below consumes them
// one dependent load at a time, which would otherwise cost one DRAM round trip per entry
__device__ __forceinline__ void band_prefetch_row(const HvbDev& D, int r, const cplx* __restrict__ Vt, long long T) {
    const int lo = __ldg(D.row_ptr + r), hi = __ldg(D.row_ptr + r + 1);
    for (int e = lo; e < hi; ++e) {
        const int vid = (__ldg(D.row_ent + e) >> 1) & 0x7FFFF;
        if (vid >= D.nvc) asm volatile("prefetch.global.L1 [%0];" ::"l"(Vt + (long long)(vid - D.nvc) * T));
    }
}

// assemble system row r into a window row whose first column is k0 (row entries are sorted by column)
template <int B, int PP>
__device__ __forceinline__ void band_assemble_row(const HvbDev& D, int r, int k0, const cplx* __restrict__ Vt, long long T,
                                                  BandRow<B, PP>& row) {
    int e = __ldg(D.row_ptr + r);
    const int hi = __ldg(D.row_ptr + r + 1);
    int word = e < hi ? __ldg(D.row_ent + e) : 0;
    int ecol = e < hi ? (int)(((unsigned)word >> 20) & 0x7FFu) : -1;
    auto cell = [&](int col) {
        cplx acc = cmk(0.0, 0.0);
        while (ecol == col) {
            const int vid = (word >> 1) & 0x7FFFF;
            const cplx v = vid < D.nvc ? D.const_vals[vid] : Vt[(long long)(vid - D.nvc) * T];
            if (word & 1) { acc.x -= v.x; acc.y -= v.y; } else { acc.x += v.x; acc.y += v.y; }
            ++e;
            word = e < hi ? __ldg(D.row_ent + e) : 0;
            ecol = e < hi ? (int)(((unsigned)word >> 20) & 0x7FFu) : -1;
        }
        return acc;
    };
#pragma unroll
    for (int c = 0; c <= 2 * B; ++c) row.a[c] = cell(k0 + c < D.n ? k0 + c : -2);
#pragma unroll
    for (int i = 0; i < PP; ++i) row.y[i] = cell(i < D.P ? D.n + i : -2);
}

template <int B, int PP>
__global__ void __launch_bounds__(HVB_BAND_THREADS) hvb_bandf_kernel(const HvbDev D) {
    constexpr int WD = 2 * B + 1;
    const int n = D.n, P = D.P;
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // finished pivot rows are a sequential stream of records [inv(u_kk), u_k,k+1 .. u_k,k+2B, y_k0 .. y_k,P-1]:
    // written front to back by the elimination, read back to front by the substitution
    const int RS = WD + P;
    cplx* __restrict__ Rt = D.W + t;                                   // element e of record k at Rt[(k*RS + e)*T]
    cplx* __restrict__ Xs = Rt + (long long)n * RS * T;               // x(slot, r) of the rows the epilogue reads
    cplx* __restrict__ Vt = Xs + (long long)D.band_nslots * P * T;    // values
    const long long total = D.nS * D.nF;
    unsigned bad = 0;

    for (long long q = t; q < total; q += T) {
        const long long s = q / D.nF, fi = q - s * D.nF;
        const double f = __ldg(D.f + fi);
        const double* __restrict__ srow = D.samples + s * D.nR;

        // ---- value program ----------------------------------------------------------------------
        {
            const double w = __dmul_rn(HVB_TWOPI, f);
            for (int o = 0; o < D.nfast; ++o) {
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
                eval_simple_op(D, op, f, fi, srow, tmp - out);
                const int cnt = __ldg(op) == OP_TLINE ? 9 : 1;
                for (int e = 0; e < cnt; ++e) Vt[(long long)(out + e) * T] = tmp[e];
            }
        }

        // ---- elimination: window rows k .. k+B in registers --------------------------------------
        BandRow<B, PP> R[B + 1];
#pragma unroll
        for (int i = 0; i < B; ++i) {
            if (i < n) band_assemble_row<B, PP>(D, i, 0, Vt, T, R[i]);
            else {
#pragma unroll
                for (int c = 0; c < WD; ++c) R[i].a[c] = cmk(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < PP; ++r) R[i].y[r] = cmk(0.0, 0.0);
            }
        }
        double pmin = 1e300, pmax = 0.0;
        cplx* __restrict__ wp = Rt;
        for (int r = B; r < B + HVB_BAND_PREFETCH && r < n; ++r) band_prefetch_row(D, r, Vt, T);
        for (int k = 0; k < n; ++k) {
            if (k + B + HVB_BAND_PREFETCH < n) band_prefetch_row(D, k + B + HVB_BAND_PREFETCH, Vt, T);
            if (k + B < n) band_assemble_row<B, PP>(D, k + B, k, Vt, T, R[B]);
            else {
#pragma unroll
                for (int c = 0; c < WD; ++c) R[B].a[c] = cmk(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < PP; ++r) R[B].y[r] = cmk(0.0, 0.0);
            }
            int p = 0;
            double best = fabs(R[0].a[0].x) + fabs(R[0].a[0].y);
            if (!(best >= 0.0)) best = -1.0;                 // NaN never wins
#pragma unroll
            for (int i = 1; i <= B; ++i) {
                const double m = fabs(R[i].a[0].x) + fabs(R[i].a[0].y);
                if (m > best) { best = m; p = i; }
            }
            pmin = fmin(pmin, best);
            pmax = fmax(pmax, best);
#pragma unroll
            for (int i = 1; i <= B; ++i) {
                if (p == i) {
#pragma unroll
                    for (int c = 0; c < WD; ++c) { const cplx x = R[0].a[c]; R[0].a[c] = R[i].a[c]; R[i].a[c] = x; }
#pragma unroll
                    for (int r = 0; r < PP; ++r) { const cplx x = R[0].y[r]; R[0].y[r] = R[i].y[r]; R[i].y[r] = x; }
                }
            }
            const cplx inv = crecip_fast(R[0].a[0]);
            *wp = inv; wp += T;
#pragma unroll
            for (int c = 1; c < WD; ++c) { *wp = R[0].a[c]; wp += T; }
#pragma unroll
            for (int r = 0; r < PP; ++r) if (r < P) { *wp = R[0].y[r]; wp += T; }
#pragma unroll
            for (int i = 1; i <= B; ++i) {
                const cplx l = cmul(R[i].a[0], inv);
#pragma unroll
                for (int c = 1; c < WD; ++c) cfnma(R[i].a[c], l, R[0].a[c]);
#pragma unroll
                for (int r = 0; r < PP; ++r) cfnma(R[i].y[r], l, R[0].y[r]);
            }
            // slide: row i becomes row i-1, column c becomes column c-1
#pragma unroll
            for (int i = 1; i <= B; ++i) {
#pragma unroll
                for (int c = 1; c < WD; ++c) R[i - 1].a[c - 1] = R[i].a[c];
                R[i - 1].a[WD - 1] = cmk(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < PP; ++r) R[i - 1].y[r] = R[i].y[r];
            }
        }
        if (!(pmin >= 2.2250738585072014e-308)) bad |= HVB_STATUS_ZERO_PIVOT_BIT;
        if (!(pmax <= 1.7976931348623157e308)) bad |= HVB_STATUS_NONFINITE_BIT;

        // ---- back substitution: the last 2B solution rows slide through registers ------------------
        {
            cplx X[2 * B][PP];
#pragma unroll
            for (int j = 0; j < 2 * B; ++j)
#pragma unroll
                for (int r = 0; r < PP; ++r) X[j][r] = cmk(0.0, 0.0);
            const long long rec = (long long)RS * T;
            const cplx* __restrict__ rp = Rt + (long long)(n - 1) * rec;
            for (int d = 1; d <= HVB_BAND_PREFETCH_BACK && n - 1 - d >= D.band_min_row; ++d)
                for (int e = 0; e < RS; ++e) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp - d * rec + e * T));
            for (int k = n - 1; k >= D.band_min_row; --k, rp -= rec) {
                if (k - HVB_BAND_PREFETCH_BACK - 1 >= D.band_min_row) {
                    const cplx* __restrict__ pf = rp - (HVB_BAND_PREFETCH_BACK + 1) * rec;
                    for (int e = 0; e < RS; ++e) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + e * T));
                }
                cplx u[WD];
#pragma unroll
                for (int c = 0; c < WD; ++c) u[c] = rp[c * T];
                cplx x[PP];
#pragma unroll
                for (int r = 0; r < PP; ++r) {
                    cplx acc = (r < P) ? rp[(WD + r) * T] : cmk(0.0, 0.0);
#pragma unroll
                    for (int j = 1; j <= 2 * B; ++j) cfnma(acc, u[j], X[j - 1][r]);
                    x[r] = cmul(acc, u[0]);
                }
                const int slot = __ldg(D.band_slot_of_row + k);
                if (slot >= 0) {
#pragma unroll
                    for (int r = 0; r < PP; ++r) if (r < P) Xs[((long long)slot * P + r) * T] = x[r];
                }
#pragma unroll
                for (int j = 2 * B - 1; j >= 1; --j)
#pragma unroll
                    for (int r = 0; r < PP; ++r) X[j][r] = X[j - 1][r];
#pragma unroll
                for (int r = 0; r < PP; ++r) X[0][r] = x[r];
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

const void* hvb_bandf_kernel_ptr(int b, int pp) {
#define HVB_BANDF(BB, PPV) if (b == BB && pp == PPV) return (const void*)hvb_bandf_kernel<BB, PPV>;
    HVB_BANDF(1, 2) HVB_BANDF(1, 4) HVB_BANDF(1, 8) HVB_BANDF(1, 16)
    HVB_BANDF(2, 2) HVB_BANDF(2, 4) HVB_BANDF(2, 8)
#undef HVB_BANDF
    return nullptr;
}
size_t hvb_bandf_words_per_thread(int b, int n, int p, int nslots, int n_dyn) {
    return (size_t)n * (size_t)(2 * b + 1 + p) + (size_t)nslots * p + (size_t)(n_dyn > 0 ? n_dyn : 1);
}

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
