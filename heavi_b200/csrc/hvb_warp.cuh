// hvb_warp.cuh -- kernel A: systems with n <= 32 unknowns, G lanes per system, rows in registers.
//
// One system (one (sample, frequency) point) is owned by a group of G in {4,8,16,32} lanes, so a
// warp works on 32/G points at once.  Lane r of the group holds row r of the augmented matrix
// [A | B] in registers: N complex matrix columns + P complex right-hand sides, N and P template
// constants (the host pads the plan to an instantiated shape with identity rows / zero columns,
// so the elimination loops carry no bounds checks at all).  Per point:
//
//   1. value program  -> per-point stamp values in shared memory (group-cooperative N-port ops,
//                        then the simple ops distributed over the lanes)
//   2. assembly       -> each lane walks its row's entry list (sorted by column, component order
//                        inside a cell = the reference's order of `+=`), accumulates a cell in
//                        registers and drops it into its shared-memory row; the row is then read
//                        back with compile-time column indices
//   3. Gauss-Jordan   -> N pivot steps with implicit partial pivoting: rows never move; the pivot
//                        lane is found with a 32-bit key butterfly (|re|+|im| like LAPACK izamax,
//                        ties to the lowest row), publishes its row through a double-buffered
//                        shared-memory line, and every other row eliminates column k.  Because
//                        rows above and below are eliminated there is no back substitution: after
//                        N steps row r holds x[col(r)] = b_r / pivot_r for all P excitations.
//   4. epilogue       -> S[j,i] from port voltages (solvers/spfn.py:159-176), stored with
//                        frequency fastest so neighbouring groups/warps write neighbouring words.
//
// One __syncthreads at kernel start (plan tables staged into shared memory); after that a group
// only ever synchronises with its own warp.
#pragma once
#include "hvb_device.cuh"

#include <type_traits>

#define HVB_WARP_THREADS 128

// compile-time loop: indices are template constants, so register arrays are always statically indexed
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& fn) {
    if constexpr (I < N) {
        fn(std::integral_constant<int, I>{});
        static_for<I + 1, N>(fn);
    }
}

__host__ __device__ constexpr int hvb_group_for(int n) { return n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }

// registers per thread are capped so that small systems keep many warps resident
template <int N, int P>
struct WarpCfg {
#ifndef HVB_THREE_BLOCK_REGS
#define HVB_THREE_BLOCK_REGS 180
#endif
#ifndef HVB_REG_FLOOR
#define HVB_REG_FLOOR 88
#endif
    static constexpr int kRegsWanted = (4 * (N + P) + 48) < HVB_REG_FLOOR ? HVB_REG_FLOOR : (4 * (N + P) + 48);
#ifdef HVB_MINB_SMALL
    static constexpr int kMinBlocks = (N <= 8 && P <= 4) ? HVB_MINB_SMALL :
#else
    static constexpr int kMinBlocks =
#endif
                                      kRegsWanted <= 64 ? 8 : kRegsWanted <= 72 ? 7 : kRegsWanted <= 80 ? 6
                                    : kRegsWanted <= 96 ? 5 : kRegsWanted <= 128 ? 4 : kRegsWanted <= 168 ? 3
                                    : kRegsWanted <= HVB_THREE_BLOCK_REGS ? 3 : kRegsWanted <= 255 ? 2 : 1;
};

template <int G>
__device__ __forceinline__ unsigned group_max_u32(unsigned key) {
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, off));
    return key;
}

// pivot key: high word of |re|+|im| with the low 5 bits replaced by (31 - lane) so that one
// integer max yields both the largest magnitude (to 15 mantissa bits) and its lowest-index owner.
__device__ __forceinline__ unsigned pivot_key(cplx v, bool excluded, int gl) {
    const double mag = fabs(v.x) + fabs(v.y);
    const unsigned hi = (unsigned)__double2hiint(mag);
    return excluded ? 0u : (((hi & ~31u) | (unsigned)(31 - gl)) + 32u);
}
// smallest / largest winning key of a point -> status bits (zero pivot, NaN/Inf)
__device__ __forceinline__ unsigned pivot_status(unsigned kmin, unsigned kmax) {
    unsigned bad = 0;
    if (kmin < 64u) bad |= HVB_STATUS_ZERO_PIVOT_BIT;               // |pivot| below 2^-1022
    if (((kmax - 32u) & ~31u) >= 0x7FF00000u) bad |= HVB_STATUS_NONFINITE_BIT;
    return bad;
}

// ------------------------------------------------------------------------------------------------
// cooperative N-port value op, executed by all G lanes of a group (base/nport.py:28-39):
//   S_ij(f) -> Y = (1/Z0) (I - S) (I + S)^-1 -> (P+1)x(P+1) indefinite block with the reference
//   node's row / column.  The inverse is applied as a solve of (I+S)^T X^T = (I-S)^T with the
//   same implicit-pivot Gauss-Jordan, rows in the group's shared-memory scratch.
// Kept out of line: it is only reached by circuits that contain an N-port block.
// ------------------------------------------------------------------------------------------------
struct CoopCtx {
    ParamView pv;
    int nvc;
    double f;
    long long fi;
    const double* srow;
    cplx* dyn;
    cplx* scratch;
};

// complex array in shared memory addressed through the shared window (an out-of-line function only
// sees generic pointers; explicit ld.shared / st.shared keeps these accesses on the LDS/STS path)
struct SmemCplx {
    unsigned base;
    __device__ __forceinline__ explicit SmemCplx(const cplx* p) : base((unsigned)__cvta_generic_to_shared(p)) {}
    __device__ __forceinline__ cplx ld(int i) const {
        cplx v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + 16u * (unsigned)i));
        return v;
    }
    __device__ __forceinline__ void st(int i, cplx v) const {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base + 16u * (unsigned)i), "d"(v.x), "d"(v.y) : "memory");
    }
};

template <int G>
static __device__ __noinline__ unsigned coop_nport(const CoopCtx* ctx, const int* __restrict__ op, int gl) {
    const ParamView& D = ctx->pv;
    const double f = ctx->f;
    const long long fi = ctx->fi;
    const double* __restrict__ srow = ctx->srow;
    const SmemCplx dyn(ctx->dyn), W(ctx->scratch);
    const int code = op[0], out = op[1] - ctx->nvc, Pn = op[2], first = op[3];
    const int ldo = Pn + 1;
    // lanes form a (row, column-slice) grid: PR row slots (power of two >= Pn), LPR lanes per row
    int PR = 1;
    while (PR < Pn) PR <<= 1;
    const int LPR = G / PR > 0 ? G / PR : 1;
    const int r = gl % PR, cg = gl / PR;
    const bool row_live = r < Pn;
    unsigned kmin = 0xffffffffu, kmax = 0u;
    if (code == COOP_NPORT_Y) {
        for (int e = gl; e < Pn * Pn; e += G)
            dyn.st(out + (e / Pn) * ldo + (e % Pn), eval_param(D, first + e, f, fi, srow));
        __syncwarp();
    } else {
        const int ld = 2 * Pn;
        SplineBasis sb;
        sb.t_off = -1; sb.nt = 0; sb.k = 0; sb.ell = 0;
        for (int e = gl; e < Pn * Pn; e += G) {
            const int i = e / Pn, j = e % Pn;
            cplx s;
            const int2 pr = D.prefs[first + e];
            if (pr.x == PK_SPLINE) {
                // all traces of a Touchstone block share one knot vector: interval search and basis
                // values are done once per lane, each further trace is a 4-term dot product
                const double2 rng = D.spl_range[pr.y];
                if (f < rng.x) s = D.spl_fill[2 * pr.y];
                else if (f > rng.y) s = D.spl_fill[2 * pr.y + 1];
                else {
                    const int4 d4 = D.spl_trace[pr.y];
                    if (d4.x != sb.t_off || d4.y != sb.nt || d4.w != sb.k) spline_basis(D.spl_t, d4.x, d4.y, d4.w, f, &sb);
                    s = spline_dot(D.spl_c + d4.z, sb);
                }
            } else {
                s = eval_param(D, first + e, f, fi, srow);
            }
            const double d = (i == j) ? 1.0 : 0.0;
            W.st(j * ld + i, cmk(d + s.x, s.y));              // (I+S)^T
            W.st(j * ld + Pn + i, cmk(d - s.x, -s.y));        // (I-S)^T
        }
        __syncwarp();
        bool done = !row_live;
        int mycol = 0;
        for (int k = 0; k < Pn; ++k) {
            const cplx mine = row_live ? W.ld(r * ld + k) : cmk(0.0, 0.0);
            const unsigned key = group_max_u32<G>(pivot_key(mine, done, r));
            kmin = min(kmin, key); kmax = max(kmax, key);
            const int p = 31 - (int)(key & 31u);
            if (r == p) { done = true; mycol = k; }
            const cplx inv = crecip_fast(W.ld(p * ld + k));
            if (row_live && r != p) {
                const cplx m = cmul(mine, inv);
                for (int c = k + 1 + cg; c < ld; c += LPR) {
                    cplx v = W.ld(r * ld + c);
                    cfnma(v, m, W.ld(p * ld + c));
                    W.st(r * ld + c, v);
                }
            }
            __syncwarp();
        }
        if (row_live) {
            const cplx inv = crecip_fast(W.ld(r * ld + mycol));
            const cplx iz = crecip_np(D.consts[op[4]]);            // (1/Z0), nport.py:32
            for (int c = cg; c < Pn; c += LPR) {
                const cplx x = cmul(W.ld(r * ld + Pn + c), inv);   // X^T[mycol][c] = X[c][mycol]
                dyn.st(out + c * ldo + mycol, (iz.y == 0.0) ? cscale(x, iz.x) : cmul(iz, x));
            }
        }
        __syncwarp();
    }
    // reference-node row / column: -row sums, -column sums, +total (nport.py:35-38), sums in index order
    for (int t = gl; t < 2 * Pn; t += G) {
        const int i = t >> 1;
        cplx acc = cmk(0.0, 0.0);
        if (t & 1) {
            for (int j = 0; j < Pn; ++j) acc = cadd(acc, dyn.ld(out + j * ldo + i));
            dyn.st(out + Pn * ldo + i, cneg(acc));
        } else {
            for (int j = 0; j < Pn; ++j) acc = cadd(acc, dyn.ld(out + i * ldo + j));
            dyn.st(out + i * ldo + Pn, cneg(acc));
            W.st(i, acc);
        }
    }
    __syncwarp();
    if (gl == 0) {
        cplx tot = cmk(0.0, 0.0);
        for (int i = 0; i < Pn; ++i) tot = cadd(tot, W.ld(i));
        dyn.st(out + Pn * ldo + Pn, tot);
    }
    __syncwarp();
    return (code == COOP_NPORT_S) ? pivot_status(kmin, kmax) : 0u;
}

// complex words taken by the block-shared plan tables (same formula on host and device)
__host__ __device__ constexpr size_t hvb_warp_shared_words_dev(int N, int nvc, int nnz_rows, int nfast, int P) {
    return (size_t)nvc + (size_t)((nfast + 1) / 2) + (size_t)((P * P + 1) / 2) +
           (size_t)((N + 1 + nnz_rows + 2 * nfast + N + 3) / 4);
}

// row entry: bit 0 negate, bits 1..19 value id, bits 20..30 column, bit 31 last entry of its cell
#define HVB_ENT_VID(e) (((e) >> 1) & 0x7FFFF)
#define HVB_ENT_COL(e) (((e) >> 20) & 0x7FF)

// ------------------------------------------------------------------------------------------------
// kernel A
// ------------------------------------------------------------------------------------------------
template <int N, int P>
__global__ void __launch_bounds__(HVB_WARP_THREADS, WarpCfg<N, P>::kMinBlocks) hvb_warp_kernel(const HvbDev D) {
    extern __shared__ __align__(16) unsigned char hvb_smem[];
    constexpr int G = hvb_group_for(N);
    constexpr int SPW = 32 / G;
    constexpr int NC = N + P;                            // columns of one augmented row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, grp = lane / G;
    const int sys_in_block = warp * SPW + grp;
    constexpr int sys_per_block = (HVB_WARP_THREADS / 32) * SPW;
    const int Pr = D.P;                                  // real ports (<= P)

    // ---- block-shared copies of the small plan tables ------------------------------------------
    const int nfast = D.nfast;
    cplx* cvS = reinterpret_cast<cplx*>(hvb_smem);                       // [nvc]
    double* fargS = reinterpret_cast<double*>(cvS + D.nvc);              // [nfast] (padded to even)
    double* escS = fargS + ((nfast + 1) & ~1);                           // [Pr*Pr] (padded to even)
    int* rowptrS = reinterpret_cast<int*>(escS + ((Pr * Pr + 1) & ~1));  // [N+1]
    int* rowentS = rowptrS + (N + 1);                                    // [nnz_rows]
    int* fcodeS = rowentS + D.nnz_rows;                                  // [nfast]
    int* foutS = fcodeS + nfast;                                         // [nfast]
    int* eporS = foutS + nfast;                                          // [N]
    const int shared_words = (int)hvb_warp_shared_words_dev(N, D.nvc, D.nnz_rows, nfast, Pr);
    for (int i = threadIdx.x; i < D.nvc; i += HVB_WARP_THREADS) cvS[i] = D.const_vals[i];
    for (int i = threadIdx.x; i < nfast; i += HVB_WARP_THREADS) {
        fargS[i] = D.fop_arg[i]; fcodeS[i] = D.fop_code[i]; foutS[i] = D.fop_out[i];
    }
    for (int i = threadIdx.x; i < Pr * Pr; i += HVB_WARP_THREADS) escS[i] = D.epi_scale[i];
    for (int i = threadIdx.x; i <= N; i += HVB_WARP_THREADS) rowptrS[i] = D.row_ptr[i];
    for (int i = threadIdx.x; i < D.nnz_rows; i += HVB_WARP_THREADS) rowentS[i] = D.row_ent[i];
    for (int i = threadIdx.x; i < N; i += HVB_WARP_THREADS) eporS[i] = D.epi_port_of_row[i];
    __syncthreads();

    cplx* regA = reinterpret_cast<cplx*>(hvb_smem) + shared_words + (size_t)sys_in_block * (D.region_a + D.region_b);
    cplx* regB = regA + D.region_a;
    const long long total = D.nS * D.nF;
    const int r = gl;
    const int nvc = D.nvc;

    for (long long q0 = (long long)blockIdx.x * sys_per_block; q0 < total;
         q0 += (long long)gridDim.x * sys_per_block) {
        const long long q = q0 + sys_in_block;
        const bool active = q < total;
        const long long qq = active ? q : total - 1;     // idle groups shadow the last point
        const long long s = qq / D.nF, fi = qq - s * D.nF;
        const double f = __ldg(D.f + fi);
        if (D.f32 && active && s == 0 && gl == 0) D.f32[fi] = (float)f;
        const double* __restrict__ srow = D.samples + s * D.nR;
        unsigned bad = 0;

        // ---- 1. value program ----------------------------------------------------------------
        if (D.ncoops > 0) {
            CoopCtx ctx{D.pv, D.nvc, f, fi, srow, regA, regB};
            for (int o = 0; o < D.ncoops; ++o) bad |= coop_nport<G>(&ctx, D.coops + o * HVB_COOP_WORDS, gl);
        }
        eval_fast_ops(fcodeS, foutS, fargS, nfast, f, srow, regA, gl, G);
        for (int o = gl; o < D.nops; o += G) eval_simple_op(D, D.ops + o * HVB_OP_WORDS, f, fi, srow, regA);
        __syncwarp();

        // ---- 2. assemble this lane's row -------------------------------------------------------
        cplx a[N], b[P];
        {
            cplx* __restrict__ rowbuf = regB + (r < N ? r : 0) * NC;
            if (r < N) {
                static_for<0, NC>([&](auto cc) { rowbuf[decltype(cc)::value] = cmk(0.0, 0.0); });
                const int lo = rowptrS[r], hi = rowptrS[r + 1];
                cplx acc = cmk(0.0, 0.0);
                for (int e = lo; e < hi; ++e) {
                    const int ent = rowentS[e];
                    const int vid = HVB_ENT_VID(ent);
                    const cplx v = (vid < nvc) ? cvS[vid] : regA[vid - nvc];
                    if (ent & 1) { acc.x -= v.x; acc.y -= v.y; } else { acc.x += v.x; acc.y += v.y; }
                    if (ent < 0) { rowbuf[HVB_ENT_COL(ent)] = acc; acc = cmk(0.0, 0.0); }
                }
            }
            static_for<0, N>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                a[c] = (r < N) ? rowbuf[c] : cmk(0.0, 0.0);
            });
            static_for<0, P>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                b[i] = (r < N) ? rowbuf[N + i] : cmk(0.0, 0.0);
            });
        }
        __syncwarp();                                     // row buffers alias the broadcast lines

        // ---- 3. Gauss-Jordan with implicit partial pivoting ------------------------------------
        bool done = r >= N;
        int mycol = 0;
        cplx myinv = cmk(0.0, 0.0);
        unsigned kmin = 0xffffffffu, kmax = 0u;
        static_for<0, N>([&](auto kk) {
            constexpr int k = decltype(kk)::value;
            const unsigned key = group_max_u32<G>(pivot_key(a[k], done, gl));
            kmin = min(kmin, key); kmax = max(kmax, key);
            const bool is_p = (gl == 31 - (int)(key & 31u));
            cplx* __restrict__ brow = regB + (k & 1) * NC;
            if (is_p) {
                static_for<k, N>([&](auto cc) { brow[decltype(cc)::value] = a[decltype(cc)::value]; });
                static_for<0, P>([&](auto ii) { brow[N + decltype(ii)::value] = b[decltype(ii)::value]; });
                done = true;
                mycol = k;
            }
            __syncwarp();
            const cplx inv = crecip_fast(brow[k]);
            cplx m = cmul(a[k], inv);
            if (is_p) { myinv = inv; m = cmk(0.0, 0.0); }
            static_for<k + 1, N>([&](auto cc) { cfnma(a[decltype(cc)::value], m, brow[decltype(cc)::value]); });
            static_for<0, P>([&](auto ii) { cfnma(b[decltype(ii)::value], m, brow[N + decltype(ii)::value]); });
        });
        bad |= pivot_status(kmin, kmax);

        // ---- 4. epilogue --------------------------------------------------------------------------
        if (D.epi_simple) {
            // Norton plan, ground-referenced ports, constant real Z0: spfn.py:173-176 reduces to
            // S_ji = (2 V_sig_j - delta_ij) sqrt|Zi| / sqrt|Zj| (denominator == V_int - V_gnd == 1 V);
            // the lane that owns the solution row of port j's signal node writes S[j, :] directly.
            const int pj = (r < N) ? eporS[mycol] : -1;
            if (pj >= 0) {
                static_for<0, P>([&](auto ii) {
                    constexpr int i = decltype(ii)::value;
                    if (i < Pr) {
                        cplx x = cmul(b[i], myinv);
                        x.x = fma(2.0, x.x, (i == pj) ? -1.0 : 0.0);
                        x.y = 2.0 * x.y;
                        const cplx sv = cscale(x, escS[pj * Pr + i]);
                        if (!(isfinite(sv.x) && isfinite(sv.y))) bad |= HVB_STATUS_NONFINITE_BIT;
                        if (active) D.S[((s * Pr + pj) * Pr + i) * D.ldS + fi] = sv;
                    }
                });
            }
        } else {
            __syncwarp();
            if (r < N) {
                static_for<0, P>([&](auto ii) {
                    constexpr int i = decltype(ii)::value;
                    regA[mycol * P + i] = cmul(b[i], myinv);
                });
            }
            __syncwarp();
            for (int e = gl; e < Pr * Pr; e += G) {
                const int j = e / Pr, i = e - j * Pr;
                const cplx Zi = eval_param(D.pv, __ldg(D.port_zpref + i), f, fi, srow);
                const cplx Zj = eval_param(D.pv, __ldg(D.port_zpref + j), f, fi, srow);
                const cplx sv = s_entry(D, regA, j, i, Zi, Zj);
                if (!(isfinite(sv.x) && isfinite(sv.y))) bad |= HVB_STATUS_NONFINITE_BIT;
                if (active) D.S[((s * Pr + j) * Pr + i) * D.ldS + fi] = sv;
            }
        }
        if (bad && active && D.status) atomicOr(D.status, (int)bad);
        __syncwarp();
    }
}
