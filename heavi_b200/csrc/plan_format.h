/* plan_format.h -- integer codes shared by the plan compiler (heavi_b200/plan.py, keep in sync)
 * and the kernels. */
#ifndef HVB_PLAN_FORMAT_H
#define HVB_PLAN_FORMAT_H

/* parameter reference kinds */
#define PK_CONST 0
#define PK_SAMPLE 1
#define PK_SAMPLE_NEG 2
#define PK_SAMPLE_INV 3
#define PK_TABLE 4
#define PK_SPLINE 5
#define PK_BETA_MUL 8    /* ((2 pi f) * consts[idx]) / c0, consts[idx] = sqrt(er): lib/tl.py's order (params.py LinePhase) */
#define PK_SMD 7         /* SMD part with parasitics: consts[idx..idx+3] = kind, a, b, c (params.py SMDImpedance) */
#define PK_BETA 6        /* (2 pi f / c0) * consts[idx], consts[idx] = sqrt(er) taken by the plan compiler */

/* simple value ops (8 int32 words: code, out, p0, p1, p2, p3, aux0, aux1); every p is a parameter reference.
 * OP_TLINE: p0 = Z0, p1 = beta, p2 = length */
#define OP_COPY 0
#define OP_RECIP 1
#define OP_CAP 2
#define OP_IND 3
#define OP_TLINE 4
#define HVB_OP_WORDS 8

/* cooperative value ops (8 int32 words: code, out, P, first_pref, z0_const, -, -, -) */
#define COOP_NPORT_S 0
#define COOP_NPORT_Y 1
#define HVB_COOP_WORDS 8

#define ROW_GROUND (-1)
#define ROW_VIRTUAL_INT (-2)

#define HVB_MAX_NPORT 32
#define HVB_MAX_SPLINE_K 3

#define HVB_STATUS_NONFINITE_BIT 1u
#define HVB_STATUS_ZERO_PIVOT_BIT 2u
#define HVB_STATUS_RERUN_BIT 4u       // two-sided chain kernel: a point was unsafe for the cut; the library repeats the call one-sided

#define HVB_TWOPI 6.283185307179586     /* np.float64(2*np.pi) */
#define HVB_C0 299792458.0

#endif
