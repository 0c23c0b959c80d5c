// hvb_device.cuh -- device-side plan view, complex helpers and the value program (stamp values).
//
// Everything a kernel needs to turn (frequency, Monte-Carlo sample) into the stamp values of
// every component.  Formulas and operation order follow the reference's per-class
// SP_matrix_compiler closures (src/heavi/base/*.py; exact semantics in SURVEY.md App. A); numpy's
// complex division (Smith's algorithm with a reciprocal scale) is reproduced with unfused
// multiplies/adds so that stamp values agree with the reference to the last few ulp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan_format.h"

typedef double2 cplx;

// what a parameter reference needs (kept separate so that out-of-line helpers can take a copy)
struct ParamView {
    const cplx* consts;
    const int2* prefs;
    const cplx* tables;
    long long table_ld;                     // frequency pitch of the tables
    const double* spl_t;
    const cplx* spl_c;
    const int4* spl_trace;
    const cplx* spl_fill;
    const double2* spl_range;
};

struct HvbDev {
    int n, ncols, P;                        // rows, rows + RHS columns (both possibly padded), real ports
    int PP;                                 // RHS columns = ncols - n (stride of the solution rows)
    int nnz_rows;                           // entries in row_ent
    int nfast, epi_simple;                  // fast value ops; closed-form epilogue flag
    int nvc, ndyn, nops, ncoops, nR, coop_scratch;
    int region_a, region_b;                 // per-system shared-memory regions (complex words)
    int block_main;                         // kernel B: complex words before its small per-column buffers
    long long nF, nS;                       // points of this launch: q = s*nF + fi
    long long ldS;                          // frequency pitch of the S output (elements)
    ParamView pv;
    const double* f;
    const double* samples;
    const cplx* const_vals;
    const int* ops;
    const int* coops;
    const int* cell_ptr;
    const int* cell_ent;
    const int* row_ptr;                     // per-row entry lists (kernel A's assembly)
    const int* row_ent;
    const int* fop_code;
    const int* fop_out;
    const double* fop_arg;
    const int* epi_port_of_row;
    const double* epi_scale;
    const int* port_rows;
    const int* port_zpref;
    cplx* S;
    float* f32;                             // optional: f rounded to float32, written by sample 0's points
    cplx* W;                                // dense dump / block-kernel / band-kernel workspace
    const int* band_slot_of_row;            // kernel C (fused): solution rows the S epilogue reads -> slot
    int band_nslots, band_min_row;
    int band_pfb;                           // pivot-row records prefetched (to L2) ahead of the back substitution
    int band_nsm;                           // value slots held in shared memory (the rest spill to W)
    const int* band_act;                    // [n] right-hand-side columns that can be non-zero in rows <= k+B
    long long band_rec_words;               // words of one point's pivot-row record stream
    const int* band_sched_ptr;              // [n+1] value ops to run before assembling row r
    const int* band_sched_op;               // >= 0: fast op index, < 0: ~(simple op index)
    const int* band_fslot;                  // [nfast] output slot
    const int* band_oslot;                  // [nops][9] output slots (-1: unused output)
    const int* band_ent;                    // row_ent with per-point value ids replaced by nvc + slot
    int* status;
    const int* gen_rows;                    // kernel D, rolled form: [n-1][4] assembly case, step case, itab base, dtab base
    const int* gen_itab;                    //   per-row integer constants (sample columns, parameter indices)
    const double* gen_dtab;                 //   per-row real constants (component values, plan-time stamp values)
};

// ------------------------------------------------------------------------------------------------
// complex arithmetic
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx cmk(double re, double im) { return make_double2(re, im); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cneg(cplx a) { return cmk(-a.x, -a.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return cmk(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmk(a.x * s, a.y * s); }

// fused form for the elimination inner loops (4 DFMA per complex multiply-subtract)
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmk(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ void cfnma(cplx& acc, cplx a, cplx b) {   // acc -= a*b
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}

// numpy-order complex multiply without contraction: (ar*br - ai*bi, ar*bi + ai*br)
__device__ __forceinline__ cplx cmul_np(cplx a, cplx b) {
    return cmk(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
               __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}

// numpy's CDOUBLE_divide (Smith's algorithm, reciprocal scale), unfused
__device__ __forceinline__ cplx cdiv_np(cplx a, cplx b) {
    const double abr = fabs(b.x), abi = fabs(b.y);
    if (abr >= abi) {
        if (abr == 0.0 && abi == 0.0) return cmk(a.x / abr, a.y / abi);
        const double rat = b.y / b.x;
        const double scl = 1.0 / __dadd_rn(b.x, __dmul_rn(b.y, rat));
        return cmk(__dmul_rn(__dadd_rn(a.x, __dmul_rn(a.y, rat)), scl),
                   __dmul_rn(__dsub_rn(a.y, __dmul_rn(a.x, rat)), scl));
    }
    const double rat = b.x / b.y;
    const double scl = 1.0 / __dadd_rn(b.y, __dmul_rn(b.x, rat));
    return cmk(__dmul_rn(__dadd_rn(__dmul_rn(a.x, rat), a.y), scl),
               __dmul_rn(__dsub_rn(__dmul_rn(a.y, rat), a.x), scl));
}

// 1/z for a stamp value: real division when z is real (numpy keeps real parameters real)
__device__ __forceinline__ cplx crecip_np(cplx z) {
    if (z.y == 0.0) return cmk(1.0 / z.x, 0.0);
    return cdiv_np(cmk(1.0, 0.0), z);
}

// Fast reciprocal for the elimination: hardware seed (>= 20 good bits) + two Newton steps.
// Not correctly rounded (<= ~1 ulp), which is irrelevant against kappa*eps of the solve.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
#ifdef HVB_HOST_SIM                         // tests/host_sim: the generated kernels compiled for the CPU
    r = 1.0 / x;
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ cplx crecip_fast(cplx z) {
    const double d = fma(z.x, z.x, z.y * z.y);
    const double r = rcp_fast(d);
    return cmk(z.x * r, -z.y * r);
}

// ------------------------------------------------------------------------------------------------
// B-spline trace evaluation: scipy's find_interval + _deBoor_D + dot, constant end fill
// (src/heavi/lib/touchstone.py:253-260 -> scipy.interpolate.interp1d(kind='cubic'))
// ------------------------------------------------------------------------------------------------
// knot interval + the k+1 non-zero basis values at x: identical for every trace that shares a knot
// vector (all P^2 traces of one Touchstone block do), so callers that evaluate many traces at one
// frequency compute this once and only redo the coefficient dot product per trace.
struct SplineBasis {
    int t_off, nt, k, ell;
    double h[HVB_MAX_SPLINE_K + 1];
};

static __device__ __noinline__ void spline_basis(const double* __restrict__ spl_t, int t_off, int nt, int k, double x,
                                                 SplineBasis* __restrict__ out) {
    const double* __restrict__ t = spl_t + t_off;
    const int nb = nt - k - 1;                      // number of basis functions
    // largest ell in [k, nb-1] with t[ell] <= x
    int lo = k, hi = nb - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(t + mid) <= x) lo = mid; else hi = mid - 1;
    }
    const int ell = lo;
    double h[HVB_MAX_SPLINE_K + 1], hh[HVB_MAX_SPLINE_K + 1];
    h[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= HVB_MAX_SPLINE_K; ++j) {
        if (j <= k) {
#pragma unroll
            for (int m = 0; m < HVB_MAX_SPLINE_K; ++m) if (m < j) hh[m] = h[m];
            h[0] = 0.0;
#pragma unroll
            for (int m = 1; m <= HVB_MAX_SPLINE_K; ++m) {
                if (m <= j) {
                    const double xb = __ldg(t + ell + m), xa = __ldg(t + ell + m - j);
                    if (xb == xa) { h[m] = 0.0; }
                    else {
                        const double w = hh[m - 1] / (xb - xa);
                        h[m - 1] = __dadd_rn(h[m - 1], __dmul_rn(w, xb - x));
                        h[m] = __dmul_rn(w, x - xa);
                    }
                }
            }
        }
    }
    out->t_off = t_off; out->nt = nt; out->k = k; out->ell = ell;
#pragma unroll
    for (int a = 0; a <= HVB_MAX_SPLINE_K; ++a) out->h[a] = (a <= k) ? h[a] : 0.0;
}

__device__ __forceinline__ cplx spline_dot(const cplx* __restrict__ c, const SplineBasis& b) {
    cplx out = cmk(0.0, 0.0);
#pragma unroll
    for (int a = 0; a <= HVB_MAX_SPLINE_K; ++a) {
        if (a <= b.k) {
            const cplx ca = c[b.ell - b.k + a];
            out.x = __dadd_rn(out.x, __dmul_rn(ca.x, b.h[a]));
            out.y = __dadd_rn(out.y, __dmul_rn(ca.y, b.h[a]));
        }
    }
    return out;
}

static __device__ __noinline__ cplx spline_eval(const double* __restrict__ spl_t, const cplx* __restrict__ spl_c,
                                                const int4* __restrict__ spl_trace, const cplx* __restrict__ spl_fill,
                                                const double2* __restrict__ spl_range, int tr, double x) {
    const double2 rng = spl_range[tr];
    if (x < rng.x) return spl_fill[2 * tr];
    if (x > rng.y) return spl_fill[2 * tr + 1];
    const int4 d = spl_trace[tr];                   // t_off, nt, c_off, k
    SplineBasis b;
    spline_basis(spl_t, d.x, d.y, d.w, x, &b);
    return spline_dot(spl_c + d.z, b);
}

// ------------------------------------------------------------------------------------------------
// Surface-mount part with package parasitics (heavi_b200/lib/smd.py, reference lib/smd.py:103-294),
// numpy's operation order: w = 2 pi f, zs = a + j w b, zc = 1 / (j w c), then
//   kind 0 resistor   zs * zc / (zs + zc)
//   kind 1 inductor   1 / (1 / zs + 1 / zc)
//   kind 2 capacitor  zs + zc
// and Function.__call__'s nan_to_num (NaN -> 0).
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ cplx eval_smd(const cplx* __restrict__ k, double f) {
    const int kind = (int)k[0].x;
    const double w = __dmul_rn(HVB_TWOPI, f);
    const cplx one = cmk(1.0, 0.0);
    const cplx zs = cmk(k[1].x, __dmul_rn(w, k[2].x));
    const cplx zc = cdiv_np(one, cmk(0.0, __dmul_rn(w, k[3].x)));
    cplx z;
    if (kind == 0) z = cdiv_np(cmul_np(zs, zc), cmk(__dadd_rn(zs.x, zc.x), __dadd_rn(zs.y, zc.y)));
    else if (kind == 1) {
        const cplx y1 = cdiv_np(one, zs), y2 = cdiv_np(one, zc);
        z = cdiv_np(one, cmk(__dadd_rn(y1.x, y2.x), __dadd_rn(y1.y, y2.y)));
    } else z = cmk(__dadd_rn(zs.x, zc.x), __dadd_rn(zs.y, zc.y));
    if (z.x != z.x) z.x = 0.0;
    if (z.y != z.y) z.y = 0.0;
    return z;
}

// ------------------------------------------------------------------------------------------------
// parameter reference -> complex value at (f, sample row, frequency index)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx eval_param(const ParamView& D, int pi, double f, long long fi,
                                           const double* __restrict__ srow) {
    const int2 pr = D.prefs[pi];
    switch (pr.x) {
        case PK_CONST: return D.consts[pr.y];
        case PK_SAMPLE: return cmk(srow[pr.y], 0.0);
        case PK_SAMPLE_NEG: return cmk(-srow[pr.y], 0.0);
        case PK_SAMPLE_INV: return cmk(1.0 / srow[pr.y], 0.0);
        case PK_TABLE: return D.tables[(long long)pr.y * D.table_ld + fi];
        case PK_SPLINE: return spline_eval(D.spl_t, D.spl_c, D.spl_trace, D.spl_fill, D.spl_range, pr.y, f);
        case PK_SMD: return eval_smd(D.consts + pr.y, f);
        case PK_BETA_MUL: return cmk(__dmul_rn(__dmul_rn(HVB_TWOPI, f), D.consts[pr.y].x) / HVB_C0, 0.0);
        case PK_BETA: {
            // TWOPI * f / c0 * sqrt(er)   (network.py:627-631), left to right
            const double b0 = __dmul_rn(HVB_TWOPI, f) / HVB_C0;
            const cplx s = D.consts[pr.y];                  // sqrt(er), computed by the plan compiler
            return cmk(__dmul_rn(b0, s.x), s.y == 0.0 ? 0.0 : __dmul_rn(b0, s.y));
        }
    }
    return cmk(0.0, 0.0);
}

// real-aware product: numpy promotes a real operand to (x + 0j) and multiplies; the zero imaginary
// part only ever contributes signed zeros, so a real scale is exact.
__device__ __forceinline__ cplx cmul_maybe_real(cplx a, cplx b) {
    if (a.y == 0.0) return cmk(__dmul_rn(a.x, b.x), __dmul_rn(a.x, b.y));
    if (b.y == 0.0) return cmk(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.x));
    return cmul_np(a, b);
}

// transmission line: ABCD -> Y and the indefinite 3x3 block's reference row/column (base/tl.py:30-60)
static __device__ __noinline__ void eval_tline(cplx Z0, cplx beta, double len, cplx* __restrict__ o) {
    // th = 1j*beta*length = (-bi*len, br*len)
    const double x = -beta.y * len, y = beta.x * len;
    double sy, cy;
    sincos(y, &sy, &cy);
    cplx ch, sh;                       // cosh(th), sinh(th)
    if (x == 0.0) { ch = cmk(cy, 0.0); sh = cmk(0.0, sy); }
    else {
        const double chx = cosh(x), shx = sinh(x);
        ch = cmk(chx * cy, shx * sy);
        sh = cmk(shx * cy, chx * sy);
    }
    const cplx a11 = ch, a22 = ch;
    const cplx a12 = cmul_maybe_real(Z0, sh);
    const cplx a21 = cmul_maybe_real(crecip_np(Z0), sh);
    const cplx y11 = cdiv_np(a22, a12);
    const cplx y12 = cdiv_np(cneg(csub(cmul_np(a11, a22), cmul_np(a12, a21))), a12);
    const cplx y21 = cdiv_np(cmk(-1.0, 0.0), a12);
    const cplx y22 = cdiv_np(a11, a12);
    const cplx s0 = cadd(y11, y12), s1 = cadd(y21, y22);
    o[0] = y11; o[1] = y12; o[2] = y21; o[3] = y22;
    o[4] = cneg(s0);
    o[5] = cneg(s1);
    o[6] = cneg(cadd(y11, y21));
    o[7] = cneg(cadd(y12, y22));
    o[8] = cadd(s0, s1);
}

// Same algebra with ONE complex reciprocal (1/a12) in place of the four Smith divisions and with fused
// multiplies: a few ulp away from the reference's values instead of bit-faithful.  Used by the banded
// solve kernel, where a chain of lines makes this function the largest single cost; the assembly dump
// and the other solve kernels keep eval_tline.
__device__ __forceinline__ void eval_tline_fast(cplx Z0, cplx beta, double len, cplx* __restrict__ o) {
    const double x = -beta.y * len, y = beta.x * len;
    double sy, cy;
    sincos(y, &sy, &cy);
    cplx ch, sh;
    if (x == 0.0) { ch = cmk(cy, 0.0); sh = cmk(0.0, sy); }
    else {
        const double chx = cosh(x), shx = sinh(x);
        ch = cmk(chx * cy, shx * sy);
        sh = cmk(shx * cy, chx * sy);
    }
    const cplx a12 = cmul(Z0, sh);
    const cplx a21 = cmul(crecip_fast(Z0), sh);
    const cplx r = crecip_fast(a12);
    const cplx y11 = cmul(ch, r);
    const cplx y12 = cneg(cmul(csub(cmul(ch, ch), cmul(a12, a21)), r));
    const cplx y21 = cneg(r);
    const cplx s0 = cadd(y11, y12), s1 = cadd(y21, y11);
    o[0] = y11; o[1] = y12; o[2] = y21; o[3] = y11;
    o[4] = cneg(s0);
    o[5] = cneg(s1);
    o[6] = cneg(cadd(y11, y21));
    o[7] = cneg(cadd(y12, y11));
    o[8] = cadd(s0, s1);
}

// ------------------------------------------------------------------------------------------------
// one simple value op -> writes its value(s) into dyn[] (per-point value array)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void eval_simple_op(const HvbDev& D, const int* __restrict__ op, double f,
                                               long long fi, const double* __restrict__ srow,
                                               cplx* __restrict__ dyn) {
    const int code = op[0];
    const int out = op[1] - D.nvc;
    const cplx p0 = eval_param(D.pv, op[2], f, fi, srow);
    switch (code) {
        case OP_COPY:            // base/admittance.py:38-40
            dyn[out] = p0;
            break;
        case OP_RECIP:           // base/impedance.py:28-30, base/port.py:39
            dyn[out] = crecip_np(p0);
            break;
        case OP_CAP: {           // base/capacitor.py:30-32   1j*2*pi*f*C
            const double w = __dmul_rn(HVB_TWOPI, f);
            dyn[out] = (p0.y == 0.0) ? cmk(0.0, __dmul_rn(w, p0.x))
                                     : cmk(-__dmul_rn(w, p0.y), __dmul_rn(w, p0.x));
            break;
        }
        case OP_IND: {           // base/inductor.py:31-33    1/(1j*2*pi*f*L)
            const double w = __dmul_rn(HVB_TWOPI, f);
            if (p0.y == 0.0) dyn[out] = cmk(0.0, -(1.0 / __dmul_rn(w, p0.x)));
            else dyn[out] = cdiv_np(cmk(1.0, 0.0), cmk(-__dmul_rn(w, p0.y), __dmul_rn(w, p0.x)));
            break;
        }
        case OP_TLINE: {         // base/tl.py:30-60
            const cplx beta = eval_param(D.pv, op[3], f, fi, srow);
            // the length is a parameter like any other (a sweep / Monte-Carlo / optimiser value changes it run by
            // run: base/tl.py:52 reads `self.length.value` inside every run_sp)
            eval_tline(p0, beta, eval_param(D.pv, op[4], f, fi, srow).x, dyn + out);
            break;
        }
    }
}

// fast value ops (real constant or plain sample argument), branch-free; tables may live in shared
// or global memory.  base/capacitor.py:31, inductor.py:32, impedance.py:29, admittance.py:39, port.py:39
__device__ __forceinline__ void eval_fast_ops(const int* __restrict__ fcode, const int* __restrict__ fout,
                                              const double* __restrict__ farg, int nfast, double f,
                                              const double* __restrict__ srow, cplx* __restrict__ dyn,
                                              int start, int stride) {
    const double w = __dmul_rn(HVB_TWOPI, f);
    for (int o = start; o < nfast; o += stride) {
        const int code = fcode[o];
        double p = farg[o];
        if (code & 4) p = srow[(int)p];
        const double t = (code & 2) ? __dmul_rn(w, p) : p;           // CAP / IND act on w*p
        const double v = (code & 1) ? 1.0 / t : t;                   // RECIP / IND invert
        const bool imag = (code & 2) != 0;
        dyn[fout[o]] = cmk(imag ? 0.0 : v, imag ? ((code & 1) ? -v : v) : 0.0);
    }
}

// value id -> value (constants come from global memory through the read-only path)
__device__ __forceinline__ cplx load_value(const HvbDev& D, int vid, const cplx* __restrict__ dyn) {
    if (vid < D.nvc) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(D.const_vals) + vid);
        return v;
    }
    return dyn[vid - D.nvc];
}

// gather one matrix cell from the cell program (sum in component order, like the reference's `+=`)
__device__ __forceinline__ cplx gather_cell(const HvbDev& D, int cell, const cplx* __restrict__ dyn) {
    const int lo = __ldg(D.cell_ptr + cell), hi = __ldg(D.cell_ptr + cell + 1);
    cplx acc = cmk(0.0, 0.0);
    for (int e = lo; e < hi; ++e) {
        const int ent = __ldg(D.cell_ent + e);
        const cplx v = load_value(D, ent >> 1, dyn);
        if (ent & 1) { acc.x -= v.x; acc.y -= v.y; }
        else { acc.x += v.x; acc.y += v.y; }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// S from node voltages, solvers/spfn.py:159-176.  xs[row*P + i] = voltage of system row `row` under
// excitation i.  Port rows: >=0 system row, -1 ground (0 V), -2 "reference node + source voltage".
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx port_volt(const HvbDev& D, const cplx* __restrict__ xs, int q, int which, int i) {
    const int r = __ldg(D.port_rows + 3 * q + which);
    if (r >= 0) return xs[r * D.PP + i];
    if (r == ROW_GROUND) return cmk(0.0, 0.0);
    // ROW_VIRTUAL_INT: V_int = V_ref + E, E = 1 V on the excited port and 0 V elsewhere
    const int rg = __ldg(D.port_rows + 3 * q + 2);
    cplx v = rg >= 0 ? xs[rg * D.PP + i] : cmk(0.0, 0.0);
    if (q == i) v.x += 1.0;
    return v;
}

__device__ __forceinline__ cplx cdiv_plain(cplx a, cplx b) {
    const double d = fma(b.x, b.x, b.y * b.y);
    const double r = rcp_fast(d);
    return cmk((a.x * b.x + a.y * b.y) * r, (a.y * b.x - a.x * b.y) * r);
}

__device__ __forceinline__ cplx s_entry(const HvbDev& D, const cplx* __restrict__ xs, int j, int i,
                                        cplx Zi, cplx Zj) {
    const cplx Vi1 = port_volt(D, xs, i, 0, i), Vi3 = port_volt(D, xs, i, 1, i), Vi2 = port_volt(D, xs, i, 2, i);
    const cplx Vo1 = port_volt(D, xs, j, 0, i), Vo3 = port_volt(D, xs, j, 1, i), Vo2 = port_volt(D, xs, j, 2, i);
    // numerator = (Vo1 - Vo2) + conj(Zj) * (Vo1 - Vo3) / Zj
    cplx num = cadd(csub(Vo1, Vo2), cdiv_plain(cmul(cconj(Zj), csub(Vo1, Vo3)), Zj));
    // denominator = (Vi1 - Vi2) - Zi * (Vi1 - Vi3) / Zi
    cplx den = csub(csub(Vi1, Vi2), cdiv_plain(cmul(Zi, csub(Vi1, Vi3)), Zi));
    cplx s = cdiv_plain(num, den);
    if (Zi.x == Zj.x) return s;                        // sqrt|Re Zi| / sqrt|Re Zj| == 1
    const double sc = sqrt(fabs(Zi.x)) / sqrt(fabs(Zj.x));
    return cscale(s, sc);
}
