/*
 * heavi_b200.h -- C-ABI of libheavi_b200.so, the B200 (sm_100a) engine behind heavi's
 * frequency-sweep S-parameter path.
 *
 * The reference (FennisRobert/heavi, pure Python) has no FFI layer; its seam for this path is
 *     Network.run_sp(frequencies, reinitialize=False) -> Sparameters      src/heavi/network.py:354-421
 * whose body (1) asks every component for a Python closure and lets each add a dense (k,k,nF)
 * block into a dense (N+M,N+M,nF) complex128 tensor              network.py:385-394, base/[star].py
 *             (2) builds `indices[P,4]` and `impedance_vector[M,nF]`       network.py:397-404
 *             (3) calls solve_MNA_RF_nopgb(As, Zs, indices, f, N, M)       solvers/spfn.py:106-179
 * The entry points below replace (1)+(2)+(3) in one call: the host freezes the netlist into the
 * arrays of `hvb_plan_desc` (a stamp table compiled into a value program and a per-cell gather
 * program, see heavi_b200/plan.py) and the library assembles, factors and converts every
 * (Monte-Carlo sample, frequency) point on the GPU.  Nothing here takes or returns a torch type;
 * all arguments are plain pointers and sizes.  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add to network.py to bind these.
 *
 * Conventions: every function returns 0 on success or a negative hvb_status; the message for the
 * last failure on the calling thread is hvb_last_error().  Complex numbers are interleaved
 * (re, im) float64 pairs.  A plan belongs to one CUDA device and is not thread-safe; distinct
 * plans are independent.  Caller keeps ownership of every array it passes in; the library copies
 * what it needs during the call.
 */
#ifndef HEAVI_B200_H
#define HEAVI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hvb_plan hvb_plan;

enum hvb_status {
    HVB_OK = 0,
    HVB_ERR_ARG = -1,        /* bad argument / malformed plan                                   */
    HVB_ERR_CUDA = -2,       /* CUDA runtime error (message has the CUDA string)                */
    HVB_ERR_UNSUPPORTED = -3 /* valid request this build cannot serve (e.g. N-port > 32 ports)  */
};

/* Bits of the per-run status word (hvb_run_* `status_out[0]`).  The reference raises LinAlgError
 * from numba's lstsq when the matrix holds NaN/Inf; a zero pivot is where the
 * reference silently returns the SVD minimum-norm solution (App. B-7). */
#define HVB_STATUS_NONFINITE 1
#define HVB_STATUS_ZERO_PIVOT 2
#define HVB_STATUS_RERUN 4      /* hvb_run_device only: the two-sided chain kernel met a point it cannot serve safely and the
                                   result is invalid -- hvb_plan_set_tunable(plan, "gen_two_sided", 0) and run again.  The host
                                   entries (hvb_run_host*, hvb_run_stats_host) do this themselves and never report the bit. */

/* Frozen netlist.  Field meanings follow heavi_b200/plan.py:DevicePlan one to one.
 * Replaces: the list of SP_matrix_compiler closures (network.py:385-386) plus `indices`
 * (network.py:400-404). */
typedef struct hvb_plan_desc {
    int32_t n;             /* rows of the system as laid out (incl. identity padding rows)      */
    int32_t ncols;         /* n + right-hand-side columns (== n for a dump-only plan)           */
    int32_t P;             /* ports (<= ncols - n; extra RHS columns are zero padding)          */
    int32_t n_consts;      /* length of the per-run `consts` array                              */
    int32_t n_prefs;       /* parameter references (kind, index) pairs                          */
    int32_t n_const_vals;  /* plan-time stamp values (id 0 is 1+0j)                             */
    int32_t n_dyn;         /* per-point stamp values                                            */
    int32_t n_ops;         /* simple value ops, 8 int32 words each                              */
    int32_t n_coops;       /* cooperative value ops (N-port blocks), 8 int32 words each         */
    int32_t n_samples_cols;/* columns of the per-sample parameter matrix                        */
    int32_t coop_scratch;  /* complex scratch words a cooperative op needs per point            */
    int32_t n_traces;      /* B-spline traces                                                   */
    int32_t n_knots;       /* total knots                                                       */
    int32_t n_coefs;       /* total spline coefficients                                         */
    int32_t n_row_ent;     /* entries of the per-row lists                                      */
    int32_t n_fast_ops;    /* value ops with a real-constant / plain-sample argument            */
    int32_t epi_simple;    /* 1: closed-form epilogue S_ji = (2 V_sig_j - d_ij) sqrt(Zi/Zj)     */
    int32_t kernel_hint;   /* 0 = library decides (banded kernel from 33 unknowns on when the
                              structure allows), 1 = dense kernels only, 2 = banded kernel whenever
                              the structure allows (what the host asks for on large batches of
                              narrow-banded systems with 9..32 unknowns)                        */
    const double*  const_vals; /* [n_const_vals] complex                                        */
    const int32_t* ops;        /* [n_ops*8]                                                     */
    const int32_t* coops;      /* [n_coops*8]                                                   */
    const int32_t* cell_ptr;   /* [n*ncols+1] CSR over matrix cells, row-major                  */
    const int32_t* cell_ent;   /* [cell_ptr[n*ncols]]  (value id << 1) | negate                 */
    const int32_t* row_ptr;    /* [n+1] per-row entry lists (same stamps as the cell CSR)       */
    const int32_t* row_ent;    /* [n_row_ent] bit0 negate | value id<<1 | column<<20 | last<<31 */
    const int32_t* fop_code;   /* [n_fast_ops] bits 0-1 op, bit 2: argument is a sample column  */
    const int32_t* fop_out;    /* [n_fast_ops] per-point value slot                             */
    const double*  fop_arg;    /* [n_fast_ops] constant, or sample column                       */
    const int32_t* epi_port_of_row; /* [n] port whose signal node is this row, else -1          */
    const double*  epi_scale;  /* [P*P] sqrt|Zi|/sqrt|Zj| at [j*P+i]                            */
    const int32_t* port_rows;  /* [P*3] sig,int,gnd system rows; -1 ground, -2 int = gnd + E    */
    const int32_t* port_zpref; /* [P] parameter reference of each port's Z0                     */
    const double*  spl_t;      /* [n_knots]                                                     */
    const double*  spl_c;      /* [n_coefs] complex                                             */
    const int32_t* spl_trace;  /* [n_traces*4] t_off, nt, c_off, k                              */
    const double*  spl_fill;   /* [n_traces*2] complex: value below / above the data range      */
    const double*  spl_range;  /* [n_traces*2] first / last data abscissa                       */
} hvb_plan_desc;

/* ------------------------------------------------------------------------------------------------------
 * Stamp table: the netlist as the reference's run_sp sees it (network.py:385-404), and the entry that compiles
 * it below the ABI.  One record per component in creation order: its kind, its MNA indices in the reference's
 * per-class `ids` order (ground = 0, repeated indices kept: base/port.py:23-26, base/tl.py:21-23,
 * base/nport.py:19-22, two-terminal parts [n1, n2]) and its parameters as tagged slots; plus `port_table[P][4]`
 * = exactly `indices` of network.py:400-404, the source row of every port, N and M.  No Python is needed to
 * go from this to S: hvb_plan_create_from_stamps -> hvb_run_host (consts / prefs of hvb_run_args may then be
 * NULL: the plan keeps the ones the table implies).
 * ------------------------------------------------------------------------------------------------------ */
enum hvb_comp_kind { HVB_COMP_PORT = 0, HVB_COMP_RESISTOR = 1, HVB_COMP_ADMITTANCE = 2, HVB_COMP_IMPEDANCE = 3,
                     HVB_COMP_CAPACITOR = 4, HVB_COMP_INDUCTOR = 5, HVB_COMP_TLINE = 6, HVB_COMP_NPORT_S = 7, HVB_COMP_NPORT_Y = 8 };
enum hvb_par_kind {
    HVB_PAR_CONST = 0,      /* v[0] + j v[1], fixed for the whole run (numeric.py Scalar)                            */
    HVB_PAR_SAMPLE = 1,     /* column `index` of the per-sample matrix (Random / Param / optimiser variable)         */
    HVB_PAR_SAMPLE_NEG = 2, /* minus that column (numeric.py Negative)                                               */
    HVB_PAR_SAMPLE_INV = 3, /* reciprocal of that column (numeric.py Inverse)                                        */
    HVB_PAR_TABLE = 4,      /* row `index` of hvb_run_args.tables: an arbitrary callable tabulated by the host       */
    HVB_PAR_SPLINE = 5,     /* B-spline trace `index` (Touchstone data, lib/touchstone.py:253-260)                   */
    HVB_PAR_BETA = 6,       /* (2 pi f / c0) * (v[0] + j v[1]), v = sqrt(er): Network.transmissionline, :627-631     */
    HVB_PAR_SMD = 7,        /* surface-mount part, `index` = 0 R / 1 L / 2 C, v[0..2] = its three constants          */
    HVB_PAR_BETA_MUL = 8    /* ((2 pi f) * v[0]) / c0, v[0] = sqrt(er): the operation order of lib/tl.py             */
};
typedef struct hvb_param { int32_t kind; int32_t index; double v[4]; } hvb_param;
/* parameters per component, in this order: port [Z0]; resistor [G = 1/R, const]; admittance [Y]; impedance [Z];
 * capacitor [C]; inductor [L]; line [Z0, beta, length]; N-port S [S_00 .. row-major .. S_(P-1)(P-1), Z0 const];
 * N-port Y [Y_00 ..] */
typedef struct hvb_stamp_table {
    int32_t n_components;
    const int32_t* comp_kind;   /* [n_components] hvb_comp_kind                                                */
    const int32_t* comp_port;   /* [n_components] port index (1-based, network.py:466) of a port, else 0       */
    const int32_t* ids_off;     /* [n_components+1]                                                            */
    const int32_t* ids;         /* MNA indices                                                                 */
    const int32_t* par_off;     /* [n_components+1]                                                            */
    const hvb_param* par;
    int32_t P;
    const int32_t* port_table;  /* [P*4] (iport-1, sig, int, gnd)                                              */
    const int32_t* port_src;    /* [P] source row of each port                                                 */
    int32_t N, M;
    int32_t n_traces, n_knots, n_coefs;   /* splines, laid out like hvb_plan_desc                              */
    const double* spl_t; const double* spl_c; const int32_t* spl_trace; const double* spl_fill; const double* spl_range;
} hvb_stamp_table;
enum hvb_solver_mode { HVB_MODE_AUTO = 0, HVB_MODE_FULL = 1, HVB_MODE_NORTON = 2, HVB_MODE_DUMP = 3 };
enum hvb_row_order { HVB_ORDER_AUTO = 0, HVB_ORDER_NATURAL = 1, HVB_ORDER_RCM = 2 };

/* Compile + create: picks the solver mode (Norton folding of the ports when legal), the row order and the kernel
 * like the Python front-end does -- generated chain kernel, generated hub kernel, register-resident dense kernel
 * in the shape hvb_warp_shape returns, banded kernel, dense block kernel, in that order of preference. */
int hvb_plan_create_from_stamps(const hvb_stamp_table* table, int32_t mode, int device, hvb_plan** out);

/* Host-only (no CUDA call): the compiled arrays, for inspection and for the tests that compare this compiler with
 * heavi_b200/plan.py entry by entry.  `desc_out` points into memory owned by the returned object. */
typedef struct hvb_compiled hvb_compiled;
int hvb_stamps_compile_host(const hvb_stamp_table* table, int32_t mode, int32_t pad_rows, int32_t pad_rhs, int32_t order,
                            hvb_compiled** out, hvb_plan_desc* desc_out, const double** consts_out, const int32_t** prefs_out,
                            int32_t info_out[8] /* mode, n_real, band, max_nport, n_tables, n_sample_cols, n_table_params, 0 */,
                            const int32_t** table_prefs_out, const int32_t** table_slots_out /* per table parameter: its
                            parameter reference and the constant slot a scalar-valued callable is written to */);
void hvb_compiled_free(hvb_compiled* c);

/* Per-run parameters: what only the host can evaluate (callable parameters tabulated with the
 * reference's own numpy arithmetic, Monte-Carlo draws made in the reference's RNG order).
 * Replaces: SimParam.__call__(f) inside every closure (numeric.py:127-129) and
 * MonteCarlo.iterate's per-sample mutation (numeric.py:498-503). */
typedef struct hvb_run_args {
    const double*  consts;     /* [n_consts] complex                                            */
    const int32_t* prefs;      /* [n_prefs*2]                                                   */
    const double*  f;          /* [nF] frequencies, float64                                     */
    int64_t        nF;
    const double*  samples;    /* [nS*n_samples_cols] row-major, may be NULL when no columns    */
    int64_t        nS;
    const double*  tables;     /* [n_tables*nF] complex, row-major; NULL when n_tables == 0     */
    int32_t        n_tables;
    float*         f32_out;    /* optional HOST [nF] (page-locked for full overlap): hvb_run_host also
                                  returns f rounded to float32 -- the reference's `Sparameters.f`
                                  (network.py:407,421) -- produced by the kernels and copied back tile
                                  by tile; NULL to skip.  Ignored by the other entry points.       */
    int64_t        out_ld;     /* hvb_run_host: frequency pitch of S_out in elements, 0 = nF (contiguous).
                                  A caller that owns one [nS][P][P][nF_total] array and runs frequency slab
                                  [lo, hi) of it passes S_out + lo and out_ld = nF_total: that is how several
                                  devices (hvb_run_host_multi) or several processes fill ONE result array.    */
} hvb_run_args;

/* Create / destroy.  `device` is a CUDA ordinal.  The plan copies everything in `desc`.
 * The library picks the kernel from the shape and structure of the system: register-resident dense
 * Gauss-Jordan for n <= 32 (the host pads to a shape hvb_warp_shape returns), banded LU with one
 * thread per point when the matrix block has a half bandwidth <= 8 in the row order the plan was
 * compiled with (heavi_b200/plan.py applies a reverse Cuthill-McKee ordering to large netlists),
 * blocked dense LU otherwise.  Results do not depend on the choice beyond rounding. */
int  hvb_plan_create(const hvb_plan_desc* desc, int device, hvb_plan** out);
void hvb_plan_destroy(hvb_plan* plan);

/* Tunables of the host pipeline.  Defaults come from the environment ONCE, at hvb_plan_create
 * (HVB_CHUNK_MB, HVB_MIN_SLAB, HVB_STATS_MB, HVB_BAND_TILE_WAVES, HVB_COPY2D, HVB_BAND_PFB); this call
 * changes them for one plan afterwards.  Names: "chunk_mb" (bytes of S per pipeline tile, MiB; 0 = default),
 * "min_slab" (points), "stats_mb" (HBM budget of hvb_run_stats_host, MiB), "band_tile_waves", "copy2d" (0/1),
 * "band_pfb" (records prefetched by the banded kernel; -1 = automatic), "gen_two_sided" (0: keep a long chain on
 * the one-sided generated kernel; the library sets it itself once a two-sided run refused a point).  Results never
 * depend on them beyond rounding. */
int hvb_plan_set_tunable(hvb_plan* plan, const char* name, double value);

/* S-parameters with HOST buffers: uploads f / samples / tables, runs the kernels in pipelined
 * chunks and writes S_out[nS][P][P][nF] (complex128, frequency fastest -- the layout of
 * Sparameters._S, sparam.py:49-63, with a leading sample axis).  `S_out` may be pageable or
 * pinned.  `status_out` (optional, 1 int32) receives the OR of HVB_STATUS_* bits.
 * Replaces: the whole body of Network.run_sp after index allocation (network.py:380-416). */
int hvb_run_host(hvb_plan* plan, const hvb_run_args* args, double* S_out, int32_t* status_out);

/* Non-blocking pair: hvb_run_host_async queues uploads, kernels and copies on the plan's streams and returns;
 * hvb_run_host_wait blocks until S_out / f32_out are complete and returns the status word.  One run in flight
 * per plan; distinct plans (the three filters of BASELINE configs[1], the models of an optimiser's inner loop)
 * overlap.  f / samples / consts / prefs are staged during the call; S_out, f32_out and `tables` must stay
 * alive until the wait.  Replaces several serial Network.run_sp calls (network.py:354-421). */
int hvb_run_host_async(hvb_plan* plan, const hvb_run_args* args, double* S_out);
int hvb_run_host_wait(hvb_plan* plan, int32_t* status_out);

/* The same call spread over several devices of one process: `plans[k]` are replicas of one netlist's plan
 * created on different devices (hvb_plan_create with the same descriptor).  Points are independent, so the
 * job is cut into contiguous blocks -- whole samples when nS >= nplans, else frequency slabs -- one block
 * per device, each driven by its own host thread through hvb_run_host and copied by its own device straight
 * into the caller's one S_out (page-locked for full overlap); no inter-device transfer, results bit-identical
 * to the single-device call.  status_out receives the OR over devices.   */
int hvb_run_host_multi(hvb_plan* const* plans, int32_t nplans, const hvb_run_args* args, double* S_out, int32_t* status_out);

/* Same computation with DEVICE buffers owned by the caller (e.g. torch tensors' data_ptr()):
 * ONE caller and ONE stream at a time per plan: the plan owns the device copy of consts / prefs and the
 * workspaces of the banded / block kernels, so a second call on another stream or thread must wait for the
 * first one's kernel (use distinct plans for concurrent work -- they are independent).
 * `args->f`, `args->samples`, `args->tables` and `d_S_out` are device pointers; consts/prefs stay
 * host pointers.  Work is enqueued on `stream` (a cudaStream_t, NULL = legacy default stream) and
 * the call returns without synchronising.  `d_status` (optional) is a device int32.
 * `ldS` is the frequency pitch of the output in elements (>= nF). */
int hvb_run_device(hvb_plan* plan, const hvb_run_args* args, double* d_S_out, int64_t ldS,
                   int32_t* d_status, void* stream);

/* Monte-Carlo statistics with HOST buffers: solves all nS samples like hvb_run_host but keeps S in
 * HBM and reduces over the sample axis on the device.  stats_out[6][P][P][nF] (float64) receives
 * mean and population std of |S|, of |S|^2 and of sqrt(|S|^2) in that order -- the six properties
 * of StatSparam (sparam.py:215-243).  Only 48 B per (j,i,f) cross PCIe instead of 16 B per sample.
 * Replaces: MonteCarlo.iterate / submit_S + StatSparameters.S(p1,p2) (numeric.py:498-507,
 * sparam.py:246-295) for the statistics use case. */
int hvb_run_stats_host(hvb_plan* plan, const hvb_run_args* args, double* stats_out, int32_t* status_out);

/* Assemble only: writes the dense matrix the plan describes, W[nS][nF][n][ncols] complex128, to
 * HOST memory.  With a "dump" plan (n = ncols = N+M, ground kept) this is the reference's
 * `mna_matrix[:, :, f]` (network.py:392-394), used by the parity tests. */
int hvb_assemble_host(hvb_plan* plan, const hvb_run_args* args, double* W_out);

/* Shape query: the register-resident kernel exists for a fixed list of (rows, RHS columns)
 * shapes.  Given the real system size, ports and the largest N-port block, returns the shape the
 * host should pad the plan to (identity rows, zero RHS columns), or 0/0 when the system must go
 * to the block kernel unpadded. */
int hvb_warp_shape(int32_t n, int32_t P, int32_t max_nport, int32_t* n_pad, int32_t* p_pad);

/* Introspection for the benchmark / tests. */
typedef struct hvb_kernel_info {
    int32_t kernel_class;   /* 0 = warp kernel (n <= 32), 1 = dense block kernel, 2 = banded kernel,
                               3 = generated chain kernel (see hvb_codegen_source)                */
    int32_t group;          /* warp kernel: padded rows N;  block kernel: panel width NB;
                               banded kernel: instantiated half bandwidth B                       */
    int32_t pmax;           /* warp kernel: padded RHS columns                                   */
    int32_t threads;        /* threads per CTA                                                   */
    int32_t smem_bytes;     /* dynamic shared memory per CTA                                     */
    int32_t regs;           /* registers per thread (cudaFuncGetAttributes)                      */
    int32_t ctas_per_sm;    /* occupancy                                                         */
    int32_t sm_count;
    int32_t rows_per_lane;  /* warp kernel: matrix rows held by one lane (1 or 2)                */
    int32_t stream_words;   /* banded kernel: complex words of pivot-row records one point writes
                               to global memory once and reads back once; 0 for the other classes */
    int64_t launches;       /* kernel launches issued by this plan so far                        */
    int32_t gen_rolled;     /* generated kernel: -1 not compiled yet, 0 straight-line, 1 row loop */
    int32_t gen_cases;      /* generated kernel, row loop: distinct row shapes                    */
    int32_t local_bytes;    /* generated kernel: local memory per thread (spills)                 */
    int32_t gen_scratch_acc;/* generated kernel: port-pair sums kept in the output array (P > 4)  */
    int64_t gen_flops_per_solve; /* generated kernel: executed real flops of elimination + port
                               recurrences per point (stamp values excluded)                      */
    int32_t gen_hub;        /* generated kernel: 0 chain kernel (tridiagonal netlist), 1 hub kernel
                               (one N-port S block + chains hanging off its nodes)                */
    int32_t gen_two_sided;  /* generated chain kernel, row loop: 1 when the chain is walked from both ends
                               towards a cut (about half the port-recurrence work)                */
} hvb_kernel_info;
int hvb_plan_kernel_info(hvb_plan* plan, hvb_kernel_info* out);

/* Kernel class 3, "generated": when the netlist is a chain -- the MNA matrix is tridiagonal in the plan's
 * row order -- or a hub -- one N-port S block (up to 8 ports, reference node = ground) with chains of parts
 * hanging off its nodes (a Touchstone device in per-port matching networks, BASELINE configs[2]) -- and
 * every port is ground-referenced with a constant real Z0, the library writes CUDA source for
 * exactly this netlist (stamp formulas, matrix cells and port positions as straight-line code or a loop over
 * row shapes), compiles it with NVRTC for sm_100a at the first run and launches one thread per point.  It
 * replaces the same reference lines as hvb_run_host (network.py:385-416, solvers/spfn.py:106-179).
 * The two entries below are host-only (no CUDA call, usable without a GPU): they return that source
 * and check that it compiles.  Call hvb_codegen_source with NULL buffers first to get the sizes. */
typedef struct hvb_codegen_info {
    int32_t supported;      /* 0: the plan cannot be served by a generated kernel, see `reason`   */
    int32_t rolled;
    int32_t n_cases;
    int32_t scratch_acc;
    int64_t src_len;        /* strlen of the source                                               */
    int64_t n_rows, n_itab, n_dtab;   /* elements of the row-loop tables (0 for straight-line code) */
    int64_t flops_per_solve;
    int32_t hub;            /* 1: hub kernel (N-port block + chains), 0: chain kernel             */
    int32_t threads;        /* threads per CTA the source is written for                          */
    int64_t smem_bytes;     /* dynamic shared memory per CTA                                      */
    int32_t two_sided;      /* 1: chain walked from both ends towards a cut (may set status bit 4, see hvb_run_host) */
    int32_t cut_row;        /* two-sided: last row of the front half, else -1                     */
    char reason[256];
} hvb_codegen_info;
int hvb_codegen_source(const hvb_plan_desc* desc, const int32_t* prefs, const double* consts /* [n_consts] complex or NULL:
                       with the run's constants the code carries them as literals, without it loads them */, int32_t mode, char* src_out, int64_t src_cap,
                       int32_t* rows_out, int32_t* itab_out, double* dtab_out, hvb_codegen_info* info);
int hvb_codegen_compile(const char* src, char* log_out, int64_t log_cap, int64_t* cubin_bytes);

/* FP64 FMA peak of the device, measured with a register-resident DFMA loop on `stream`
 * (TFLOP/s).  bench.py uses it as the roofline denominator because MEASURED_PEAKS.json has no
 * FP64 entry. */
int hvb_measure_dfma_peak(int device, double* tflops_out);

const char* hvb_last_error(void);
const char* hvb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HEAVI_B200_H */
