// hvb_block.cu -- kernel B: systems with n > 32 unknowns (or too many right-hand sides for
// kernel A), one CTA per (sample, frequency) point, persistent over points.
//
// The augmented matrix [A | B] of one point lives in a per-CTA workspace W[n][ld] in global
// memory that is sized to stay L2-resident (n = 207: 712 KB per CTA, 105 MB for 148 CTAs against
// 126 MB of L2).  Blocked right-looking LU with partial pivoting:
//
//   for each panel of NB columns:
//     panel (rows kb.., NB columns) -> shared memory; unblocked LU with partial pivoting there
//     row swaps applied to the trailing columns of W (L is never needed again, so not to the left)
//     U12 = L11^-1 A12 (thread per column, registers), kept in shared memory and written back
//     trailing update A22 -= L21 U12 as a register-tiled complex GEMM, operands from shared
//   blocked back substitution of the P right-hand sides, x left in the RHS columns of W
//   S epilogue (solvers/spfn.py:159-176)
//
// The value program and cell gather are the same code kernel A uses (hvb_device.cuh,
// coop_nport<32> on warp 0).
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

#define HVB_BLOCK_THREADS 256
#define HVB_BLOCK_WARPS (HVB_BLOCK_THREADS / 32)

template <int NB>
__global__ void __launch_bounds__(HVB_BLOCK_THREADS) hvb_block_kernel(const HvbDev D) {
    extern __shared__ __align__(16) unsigned char hvb_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = D.n, P = D.P, ncols = D.ncols, ld = D.ncols;
    // shared-memory carve-up (complex words unless noted)
    cplx* dyn = reinterpret_cast<cplx*>(hvb_smem);                 // region_a: values, later x rows
    cplx* scratch = dyn + D.region_a;                              // region_b: coop scratch
    cplx* panel = scratch + D.region_b;                            // [n][NB]
    cplx* u12 = panel + (size_t)n * NB;                            // [NB][ncols]
    double* redv = reinterpret_cast<double*>(u12 + (size_t)NB * ncols);   // [warps]
    int* redi = reinterpret_cast<int*>(redv + HVB_BLOCK_WARPS);    // [warps]
    int* ipiv = redi + HVB_BLOCK_WARPS;                            // [NB]
    __shared__ unsigned sh_bad;

    cplx* __restrict__ W = D.W + (size_t)blockIdx.x * n * ld;
    const long long total = D.nS * D.nF;

    for (long long q = blockIdx.x; q < total; q += gridDim.x) {
        const long long s = q / D.nF, fi = q - s * D.nF;
        const double f = __ldg(D.f + fi);
        const double* __restrict__ srow = D.samples + s * D.nR;
        unsigned bad = 0;
        if (tid == 0) sh_bad = 0;

        // ---- values -------------------------------------------------------------------------
        if (warp == 0 && D.ncoops > 0) {
            CoopCtx ctx{D.pv, D.nvc, f, fi, srow, dyn, scratch};
            for (int o = 0; o < D.ncoops; ++o) bad |= coop_nport<32>(&ctx, D.coops + o * HVB_COOP_WORDS, lane);
        }
        eval_fast_ops(D.fop_code, D.fop_out, D.fop_arg, D.nfast, f, srow, dyn, tid, HVB_BLOCK_THREADS);
        for (int o = tid; o < D.nops; o += HVB_BLOCK_THREADS)
            eval_simple_op(D, D.ops + o * HVB_OP_WORDS, f, fi, srow, dyn);
        __syncthreads();

        // ---- assemble W -----------------------------------------------------------------------
        for (int cell = tid; cell < n * ncols; cell += HVB_BLOCK_THREADS) W[cell] = gather_cell(D, cell, dyn);
        __syncthreads();

        // ---- blocked LU -----------------------------------------------------------------------
        for (int kb = 0; kb < n; kb += NB) {
            const int nb = min(NB, n - kb), m = n - kb;
            for (int idx = tid; idx < m * NB; idx += HVB_BLOCK_THREADS) {
                const int rr = idx / NB, j = idx - rr * NB;
                panel[rr * NB + j] = (j < nb) ? W[(size_t)(kb + rr) * ld + kb + j] : cmk(0.0, 0.0);
            }
            __syncthreads();
            for (int j = 0; j < nb; ++j) {
                // pivot search over rows j..m-1 of column j
                double best = -1.0;
                int bi = j;
                for (int rr = j + tid; rr < m; rr += HVB_BLOCK_THREADS) {
                    const cplx v = panel[rr * NB + j];
                    const double mag = fabs(v.x) + fabs(v.y);
                    if (mag > best || !(mag == mag)) { best = (mag == mag) ? mag : INFINITY; bi = rr; }
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (lane == 0) { redv[warp] = best; redi[warp] = bi; }
                __syncthreads();
                best = redv[0]; bi = redi[0];
#pragma unroll
                for (int w = 1; w < HVB_BLOCK_WARPS; ++w) {
                    const double ob = redv[w];
                    const int oi = redi[w];
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (!(best > 0.0)) bad |= HVB_STATUS_ZERO_PIVOT_BIT;
                if (!isfinite(best)) bad |= HVB_STATUS_NONFINITE_BIT;
                if (tid == 0) ipiv[j] = bi;
                if (bi != j && tid < NB) {
                    const cplx t = panel[j * NB + tid];
                    panel[j * NB + tid] = panel[bi * NB + tid];
                    panel[bi * NB + tid] = t;
                }
                __syncthreads();
                const cplx inv = crecip_fast(panel[j * NB + j]);
                for (int rr = j + 1 + tid; rr < m; rr += HVB_BLOCK_THREADS) {
                    const cplx l = cmul(panel[rr * NB + j], inv);
                    panel[rr * NB + j] = l;
                    for (int c = j + 1; c < nb; ++c) {
                        cplx v = panel[rr * NB + c];
                        cfnma(v, l, panel[j * NB + c]);
                        panel[rr * NB + c] = v;
                    }
                }
                __syncthreads();
            }
            const int c0 = kb + nb;                      // first trailing column
            const int ntrail = ncols - c0;
            // row swaps + U12 = L11^-1 A12, one thread per trailing column
            for (int c = tid; c < ntrail; c += HVB_BLOCK_THREADS) {
                for (int j = 0; j < nb; ++j) {
                    const int p = ipiv[j];
                    if (p != j) {
                        const cplx t = W[(size_t)(kb + j) * ld + c0 + c];
                        W[(size_t)(kb + j) * ld + c0 + c] = W[(size_t)(kb + p) * ld + c0 + c];
                        W[(size_t)(kb + p) * ld + c0 + c] = t;
                    }
                }
                cplx u[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) u[j] = (j < nb) ? W[(size_t)(kb + j) * ld + c0 + c] : cmk(0.0, 0.0);
#pragma unroll
                for (int j = 1; j < NB; ++j) {
                    if (j < nb) {                       // rows >= nb of the panel head do not exist
#pragma unroll
                        for (int i = 0; i < j; ++i) cfnma(u[j], panel[j * NB + i], u[i]);
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    if (j < nb) {
                        W[(size_t)(kb + j) * ld + c0 + c] = u[j];
                        u12[j * ncols + c] = u[j];
                    }
                }
            }
            // U11 (upper triangle of the panel head) back to W for the back substitution
            for (int idx = tid; idx < nb * nb; idx += HVB_BLOCK_THREADS) {
                const int j = idx / nb, c = idx - j * nb;
                if (c >= j) W[(size_t)(kb + j) * ld + kb + c] = panel[j * NB + c];
            }
            __syncthreads();
            // trailing update A22 -= L21 * U12, 4x4 complex register tiles, columns fastest across lanes
            const int mrows = m - nb;
            if (mrows > 0 && ntrail > 0) {
                constexpr int TR = 4, TC = 4;
                const int tiles_c = (ntrail + TC - 1) / TC, tiles_r = (mrows + TR - 1) / TR;
                for (int t = tid; t < tiles_r * tiles_c; t += HVB_BLOCK_THREADS) {
                    const int tr = t / tiles_c, tc = t - tr * tiles_c;
                    const int r0 = nb + tr * TR, cc0 = tc * TC;
                    cplx acc[TR][TC];
#pragma unroll
                    for (int i = 0; i < TR; ++i)
#pragma unroll
                        for (int j = 0; j < TC; ++j) {
                            const bool ok = (r0 + i < m) && (cc0 + j < ntrail);
                            acc[i][j] = ok ? W[(size_t)(kb + r0 + i) * ld + c0 + cc0 + j] : cmk(0.0, 0.0);
                        }
#pragma unroll 4
                    for (int k = 0; k < NB; ++k) {
                        cplx l[TR], u[TC];
#pragma unroll
                        for (int i = 0; i < TR; ++i) l[i] = (r0 + i < m) ? panel[(r0 + i) * NB + k] : cmk(0.0, 0.0);
#pragma unroll
                        for (int j = 0; j < TC; ++j) u[j] = (cc0 + j < ntrail && k < nb) ? u12[k * ncols + cc0 + j] : cmk(0.0, 0.0);
#pragma unroll
                        for (int i = 0; i < TR; ++i)
#pragma unroll
                            for (int j = 0; j < TC; ++j) cfnma(acc[i][j], l[i], u[j]);
                    }
#pragma unroll
                    for (int i = 0; i < TR; ++i)
#pragma unroll
                        for (int j = 0; j < TC; ++j)
                            if ((r0 + i < m) && (cc0 + j < ntrail)) W[(size_t)(kb + r0 + i) * ld + c0 + cc0 + j] = acc[i][j];
                }
            }
            __syncthreads();
        }

        // ---- blocked back substitution, x left in W[:, n+i] -------------------------------------
        cplx* xb = u12;                                   // [NB][P]
        const int last = ((n - 1) / NB) * NB;
        for (int kb = last; kb >= 0; kb -= NB) {
            const int nb = min(NB, n - kb);
            if (tid < P) {
                const int i = tid;
                for (int j = nb - 1; j >= 0; --j) {
                    cplx acc = W[(size_t)(kb + j) * ld + n + i];
                    for (int c = j + 1; c < nb; ++c) cfnma(acc, W[(size_t)(kb + j) * ld + kb + c], xb[c * P + i]);
                    const cplx x = cmul(acc, crecip_fast(W[(size_t)(kb + j) * ld + kb + j]));
                    xb[j * P + i] = x;
                    W[(size_t)(kb + j) * ld + n + i] = x;
                }
            }
            __syncthreads();
            for (int idx = tid; idx < kb * P; idx += HVB_BLOCK_THREADS) {
                const int rr = idx / P, i = idx - rr * P;
                cplx acc = W[(size_t)rr * ld + n + i];
                for (int j = 0; j < nb; ++j) cfnma(acc, W[(size_t)rr * ld + kb + j], xb[j * P + i]);
                W[(size_t)rr * ld + n + i] = acc;
            }
            __syncthreads();
        }

        // ---- S epilogue -------------------------------------------------------------------------
        cplx* xs = dyn;                                   // region_a >= n*P
        for (int idx = tid; idx < n * P; idx += HVB_BLOCK_THREADS) {
            const int rr = idx / P, i = idx - rr * P;
            xs[idx] = W[(size_t)rr * ld + n + i];
        }
        __syncthreads();
        for (int e = tid; e < P * P; e += HVB_BLOCK_THREADS) {
            const int j = e / P, i = e - j * P;
            const cplx Zi = eval_param(D.pv, __ldg(D.port_zpref + i), f, fi, srow);
            const cplx Zj = eval_param(D.pv, __ldg(D.port_zpref + j), f, fi, srow);
            const cplx sv = s_entry(D, xs, j, i, Zi, Zj);
            if (!(isfinite(sv.x) && isfinite(sv.y))) bad |= HVB_STATUS_NONFINITE_BIT;
            D.S[((s * P + j) * P + i) * D.ldS + fi] = sv;
        }
        if (bad) atomicOr(&sh_bad, bad);
        __syncthreads();
        if (tid == 0 && sh_bad && D.status) atomicOr(D.status, (int)sh_bad);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// dump kernel: assemble only, one warp per point, W_out[q][n][ncols]  (hvb_assemble_host)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) hvb_dump_kernel(const HvbDev D) {
    extern __shared__ __align__(16) unsigned char hvb_smem[];
    cplx* dyn = reinterpret_cast<cplx*>(hvb_smem);
    cplx* scratch = dyn + D.region_a;
    const int lane = threadIdx.x;
    const long long total = D.nS * D.nF;
    const int cells = D.n * D.ncols;
    for (long long q = blockIdx.x; q < total; q += gridDim.x) {
        const long long s = q / D.nF, fi = q - s * D.nF;
        const double f = __ldg(D.f + fi);
        const double* __restrict__ srow = D.samples + s * D.nR;
        if (D.ncoops > 0) {
            CoopCtx ctx{D.pv, D.nvc, f, fi, srow, dyn, scratch};
            for (int o = 0; o < D.ncoops; ++o) coop_nport<32>(&ctx, D.coops + o * HVB_COOP_WORDS, lane);
        }
        eval_fast_ops(D.fop_code, D.fop_out, D.fop_arg, D.nfast, f, srow, dyn, lane, 32);
        for (int o = lane; o < D.nops; o += 32) eval_simple_op(D, D.ops + o * HVB_OP_WORDS, f, fi, srow, dyn);
        __syncwarp();
        for (int cell = lane; cell < cells; cell += 32) D.W[(size_t)q * cells + cell] = gather_cell(D, cell, dyn);
        __syncwarp();
    }
}

// FP64 FMA peak probe: 8 independent chains per thread, everything in registers.
__global__ void __launch_bounds__(256) hvb_dfma_probe(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
// host-visible lookups
// ------------------------------------------------------------------------------------------------
const void* hvb_block_kernel_ptr(int nb) {
    if (nb == 8) return (const void*)hvb_block_kernel<8>;
    if (nb == 16) return (const void*)hvb_block_kernel<16>;
    return nullptr;
}
const void* hvb_dump_kernel_ptr() { return (const void*)hvb_dump_kernel; }
const void* hvb_dfma_probe_ptr() { return (const void*)hvb_dfma_probe; }
int hvb_block_threads() { return HVB_BLOCK_THREADS; }
size_t hvb_block_smem_bytes(int nb, int n, int ncols, int region_a, int region_b) {
    size_t words = (size_t)region_a + region_b + (size_t)n * nb + (size_t)nb * ncols;
    return words * sizeof(cplx) + HVB_BLOCK_WARPS * (sizeof(double) + sizeof(int)) + nb * sizeof(int) + 16;
}
