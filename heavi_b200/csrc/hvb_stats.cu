// hvb_stats.cu -- on-device Monte-Carlo statistics.
//
// The reference keeps every trial's S in a Python list and reduces on request
// (src/heavi/sparam.py:215-243, StatSparam: np.abs(S).mean(axis=0), .std(axis=0), the same for
// |S|^2 and for sqrt(|S|^2)).  For a 2-port the end-to-end cost of a batched Monte-Carlo run is the
// PCIe copy of S (64 B per solve); reducing over the sample axis on the device leaves 6 doubles per
// (j, i, f) to copy instead.  HBM-bound: each S element is read twice (mean pass, deviation pass),
// 32 B of traffic per solve-entry.  A CTA owns 32 consecutive frequencies of one (j, i) row: lane =
// frequency (every warp reads 512 contiguous bytes per sample), the 8 warps take the samples
// s = w, w+8, ... and their partial sums are combined in warp order through shared memory, so the
// result does not depend on the launch shape.
#include "hvb_device.cuh"
#include "hvb_kernels.h"

#define HVB_STATS_GROUPS 8

__global__ void __launch_bounds__(32 * HVB_STATS_GROUPS) hvb_stats_kernel(const cplx* __restrict__ S, long long nS, long long rows,
                                                                         long long nF, long long ldS, double* __restrict__ out,
                                                                         long long ldo) {
    // S[s][row][f] with pitch ldS; out[k][row][f] with pitch ldo, k = amp_mean, amp_std, pow_mean,
    // pow_std, rms_mean, rms_std
    __shared__ double part[HVB_STATS_GROUPS][3][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long long fblocks = (nF + 31) / 32;
    const long long total = rows * fblocks;
    const long long stride = rows * ldS;
    const double inv = 1.0 / (double)nS;
    for (long long blk = blockIdx.x; blk < total; blk += gridDim.x) {
        const long long row = blk / fblocks, f = (blk - row * fblocks) * 32 + lane;
        const bool live = f < nF;
        const cplx* __restrict__ p = S + row * ldS + (live ? f : 0);
        double sa = 0.0, sp = 0.0, sr = 0.0;
        if (live)
            for (long long s = g; s < nS; s += HVB_STATS_GROUPS) {
                const cplx v = p[s * stride];
                const double a = hypot(v.x, v.y);
                const double pw = a * a;
                sa += a; sp += pw; sr += sqrt(pw);
            }
        part[g][0][lane] = sa; part[g][1][lane] = sp; part[g][2][lane] = sr;
        __syncthreads();
        double ma = 0.0, mp = 0.0, mr = 0.0;
#pragma unroll
        for (int w = 0; w < HVB_STATS_GROUPS; ++w) { ma += part[w][0][lane]; mp += part[w][1][lane]; mr += part[w][2][lane]; }
        ma *= inv; mp *= inv; mr *= inv;
        __syncthreads();
        double da = 0.0, dp = 0.0, dr = 0.0;
        if (live)
            for (long long s = g; s < nS; s += HVB_STATS_GROUPS) {
                const cplx v = p[s * stride];
                const double a = hypot(v.x, v.y);
                const double pw = a * a;
                const double r = sqrt(pw);
                da += (a - ma) * (a - ma);
                dp += (pw - mp) * (pw - mp);
                dr += (r - mr) * (r - mr);
            }
        part[g][0][lane] = da; part[g][1][lane] = dp; part[g][2][lane] = dr;
        __syncthreads();
        if (g == 0 && live) {
            da = dp = dr = 0.0;
#pragma unroll
            for (int w = 0; w < HVB_STATS_GROUPS; ++w) { da += part[w][0][lane]; dp += part[w][1][lane]; dr += part[w][2][lane]; }
            double* __restrict__ o = out + row * ldo + f;
            const long long plane = rows * ldo;
            o[0 * plane] = ma;
            o[1 * plane] = sqrt(da * inv);
            o[2 * plane] = mp;
            o[3 * plane] = sqrt(dp * inv);
            o[4 * plane] = mr;
            o[5 * plane] = sqrt(dr * inv);
        }
        __syncthreads();
    }
}

int hvb_stats_threads() { return 32 * HVB_STATS_GROUPS; }
const void* hvb_stats_kernel_ptr() { return (const void*)hvb_stats_kernel; }
