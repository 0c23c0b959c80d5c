// hvb_warp2.cuh -- kernel A2: the register-resident Gauss-Jordan of hvb_warp.cuh with TWO rows per
// lane.  A system of N <= 8 unknowns then needs only 4 lanes (2 lanes for N <= 4), so a warp works
// on 8 (16) points at once: the per-step overhead that is paid once per warp instruction -- pivot
// key butterfly, publishing / re-reading the pivot row, the pivot reciprocal -- is shared by twice
// as many points, and every lane carries two independent FMA chains.  Same plan format, same
// phases, same results as kernel A; used for plans without cooperative (N-port) value ops.
#pragma once
#include "hvb_warp.cuh"

__host__ __device__ constexpr int hvb_group2_for(int n) { return n <= 4 ? 2 : n <= 8 ? 4 : n <= 16 ? 8 : 16; }

template <int N, int P>
struct Warp2Cfg {
    static constexpr int kRegsWanted = 8 * (N + P) + 56;
    static constexpr int kMinBlocks = kRegsWanted <= 96 ? 5 : kRegsWanted <= 128 ? 4 : kRegsWanted <= 168 ? 3 : 2;
};

// key with 6 row bits (rows 0..63): larger magnitude wins, ties go to the lowest row
__device__ __forceinline__ unsigned pivot_key6(cplx v, bool excluded, int row) {
    const double mag = fabs(v.x) + fabs(v.y);
    const unsigned hi = (unsigned)__double2hiint(mag);
    return excluded ? 0u : (((hi & ~63u) | (unsigned)(63 - row)) + 64u);
}

template <int N, int P>
__global__ void __launch_bounds__(HVB_WARP_THREADS, Warp2Cfg<N, P>::kMinBlocks) hvb_warp2_kernel(const HvbDev D) {
    extern __shared__ __align__(16) unsigned char hvb_smem[];
    constexpr int G = hvb_group2_for(N);
    constexpr int SPW = 32 / G;
    constexpr int NC = N + P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, grp = lane / G;
    const int sys_in_block = warp * SPW + grp;
    constexpr int sys_per_block = (HVB_WARP_THREADS / 32) * SPW;
    const int Pr = D.P;

    // ---- block-shared copies of the small plan tables (same layout as kernel A) ------------------
    const int nfast = D.nfast;
    cplx* cvS = reinterpret_cast<cplx*>(hvb_smem);
    double* fargS = reinterpret_cast<double*>(cvS + D.nvc);
    double* escS = fargS + ((nfast + 1) & ~1);
    int* rowptrS = reinterpret_cast<int*>(escS + ((Pr * Pr + 1) & ~1));
    int* rowentS = rowptrS + (N + 1);
    int* fcodeS = rowentS + D.nnz_rows;
    int* foutS = fcodeS + nfast;
    int* eporS = foutS + nfast;
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
    const int row0 = gl, row1 = gl + G;
    const bool live0 = row0 < N, live1 = row1 < N;
    const int nvc = D.nvc;

    for (long long q0 = (long long)blockIdx.x * sys_per_block; q0 < total;
         q0 += (long long)gridDim.x * sys_per_block) {
        const long long q = q0 + sys_in_block;
        const bool active = q < total;
        const long long qq = active ? q : total - 1;
        const long long s = qq / D.nF, fi = qq - s * D.nF;
        const double f = __ldg(D.f + fi);
        if (D.f32 && active && s == 0 && gl == 0) D.f32[fi] = (float)f;
        const double* __restrict__ srow = D.samples + s * D.nR;
        unsigned bad = 0;

        // ---- 1. value program (N-port blocks: only those with at most G ports are routed here) ----
        if (D.ncoops > 0) {
            CoopCtx ctx{D.pv, D.nvc, f, fi, srow, regA, regB};
            for (int o = 0; o < D.ncoops; ++o) bad |= coop_nport<G>(&ctx, D.coops + o * HVB_COOP_WORDS, gl);
        }
        eval_fast_ops(fcodeS, foutS, fargS, nfast, f, srow, regA, gl, G);
        for (int o = gl; o < D.nops; o += G) eval_simple_op(D, D.ops + o * HVB_OP_WORDS, f, fi, srow, regA);
        __syncwarp();

        // ---- 2. assemble this lane's two rows --------------------------------------------------
        cplx a0[N], a1[N], b0[P], b1[P];
        {
            auto build = [&](bool live, int row, cplx (&a)[N], cplx (&b)[P]) {
                cplx* __restrict__ rowbuf = regB + (live ? row : 0) * NC;
                if (live) {
                    static_for<0, NC>([&](auto cc) { rowbuf[decltype(cc)::value] = cmk(0.0, 0.0); });
                    const int lo = rowptrS[row], hi = rowptrS[row + 1];
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
                    a[c] = live ? rowbuf[c] : cmk(0.0, 0.0);
                });
                static_for<0, P>([&](auto ii) {
                    constexpr int i = decltype(ii)::value;
                    b[i] = live ? rowbuf[N + i] : cmk(0.0, 0.0);
                });
            };
            build(live0, row0, a0, b0);
            build(live1, row1, a1, b1);
        }
        __syncwarp();

        // ---- 3. Gauss-Jordan with implicit partial pivoting, two rows per lane -----------------
        bool done0 = !live0, done1 = !live1;
        int mycol0 = 0, mycol1 = 0;
        cplx myinv0 = cmk(0.0, 0.0), myinv1 = cmk(0.0, 0.0);
        unsigned kmin = 0xffffffffu, kmax = 0u;
        static_for<0, N>([&](auto kk) {
            constexpr int k = decltype(kk)::value;
            const unsigned key = group_max_u32<G>(max(pivot_key6(a0[k], done0, row0), pivot_key6(a1[k], done1, row1)));
            kmin = min(kmin, key); kmax = max(kmax, key);
            const int prow = 63 - (int)(key & 63u);
            const bool is0 = (row0 == prow), is1 = (row1 == prow);
            cplx* __restrict__ brow = regB + (k & 1) * NC;
            if (is0) {
                static_for<k, N>([&](auto cc) { brow[decltype(cc)::value] = a0[decltype(cc)::value]; });
                static_for<0, P>([&](auto ii) { brow[N + decltype(ii)::value] = b0[decltype(ii)::value]; });
                done0 = true; mycol0 = k;
            }
            if (is1) {
                static_for<k, N>([&](auto cc) { brow[decltype(cc)::value] = a1[decltype(cc)::value]; });
                static_for<0, P>([&](auto ii) { brow[N + decltype(ii)::value] = b1[decltype(ii)::value]; });
                done1 = true; mycol1 = k;
            }
            __syncwarp();
            const cplx inv = crecip_fast(brow[k]);
            cplx m0 = cmul(a0[k], inv), m1 = cmul(a1[k], inv);
            if (is0) { myinv0 = inv; m0 = cmk(0.0, 0.0); }
            if (is1) { myinv1 = inv; m1 = cmk(0.0, 0.0); }
            static_for<k + 1, N>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const cplx u = brow[c];
                cfnma(a0[c], m0, u);
                cfnma(a1[c], m1, u);
            });
            static_for<0, P>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                const cplx u = brow[N + i];
                cfnma(b0[i], m0, u);
                cfnma(b1[i], m1, u);
            });
        });
        {
            unsigned st = 0;
            if (kmin < 128u) st |= HVB_STATUS_ZERO_PIVOT_BIT;
            if (((kmax - 64u) & ~63u) >= 0x7FF00000u) st |= HVB_STATUS_NONFINITE_BIT;
            bad |= st;
        }

        // ---- 4. epilogue --------------------------------------------------------------------------
        if (D.epi_simple) {
            auto emit = [&](bool live, int mycol, const cplx (&b)[P], cplx myinv) {
                const int pj = live ? eporS[mycol] : -1;
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
            };
            emit(live0, mycol0, b0, myinv0);
            emit(live1, mycol1, b1, myinv1);
        } else {
            __syncwarp();
            static_for<0, P>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                if (live0) regA[mycol0 * P + i] = cmul(b0[i], myinv0);
                if (live1) regA[mycol1 * P + i] = cmul(b1[i], myinv1);
            });
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
