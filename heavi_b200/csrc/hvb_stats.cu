// hvb_stats.cu -- on-device Monte-Carlo statistics.
//
// The reference keeps every trial's S in a Python list and reduces on request
// (src/heavi/sparam.py:215-243, StatSparam: np.abs(S).mean(axis=0), .std(axis=0), the same for
// |S|^2 and for sqrt(|S|^2)).  For a 2-port the end-to-end cost of a batched Monte-Carlo run is the
// PCIe copy of S (64 B per solve); reducing over the sample axis on the device leaves 6 doubles per
// (j, i, f) to copy instead.  HBM-bound: each S element is read twice (mean pass, deviation pass),
// 32 B of traffic per solve-entry; threads map to consecutive frequencies, so every warp reads
// 512 contiguous bytes per sample.  Sums run in sample order, like numpy's axis-0 reduction.
#include "hvb_device.cuh"
#include "hvb_kernels.h"

__global__ void __launch_bounds__(256) hvb_stats_kernel(const cplx* __restrict__ S, long long nS, long long rows,
                                                        long long nF, long long ldS, double* __restrict__ out,
                                                        long long ldo) {
    // S[s][row][f] with pitch ldS; out[k][row][f] with pitch ldo, k = amp_mean, amp_std, pow_mean,
    // pow_std, rms_mean, rms_std
    const long long total = rows * nF;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / nF, f = idx - row * nF;
        const cplx* __restrict__ p = S + row * ldS + f;
        const long long stride = rows * ldS;
        double sa = 0.0, sp = 0.0, sr = 0.0;
        for (long long s = 0; s < nS; ++s) {
            const cplx v = p[s * stride];
            const double a = hypot(v.x, v.y);
            const double pw = a * a;
            sa += a; sp += pw; sr += sqrt(pw);
        }
        const double inv = 1.0 / (double)nS;
        const double ma = sa * inv, mp = sp * inv, mr = sr * inv;
        double da = 0.0, dp = 0.0, dr = 0.0;
        for (long long s = 0; s < nS; ++s) {
            const cplx v = p[s * stride];
            const double a = hypot(v.x, v.y);
            const double pw = a * a;
            const double r = sqrt(pw);
            da += (a - ma) * (a - ma);
            dp += (pw - mp) * (pw - mp);
            dr += (r - mr) * (r - mr);
        }
        double* __restrict__ o = out + row * ldo + f;
        const long long plane = rows * ldo;
        o[0 * plane] = ma;
        o[1 * plane] = sqrt(da * inv);
        o[2 * plane] = mp;
        o[3 * plane] = sqrt(dp * inv);
        o[4 * plane] = mr;
        o[5 * plane] = sqrt(dr * inv);
    }
}

const void* hvb_stats_kernel_ptr() { return (const void*)hvb_stats_kernel; }
