// hvb_warp2_s.cu -- kernel A2 (two rows per lane) instantiations for small systems.
#include "hvb_warp2.cuh"
#include "hvb_kernels.h"

const void* hvb_warp2_kernel_ptr(int n, int p) {
    if (n == 2 && p == 2) return (const void*)hvb_warp2_kernel<2, 2>;
    if (n == 3 && p == 2) return (const void*)hvb_warp2_kernel<3, 2>;
    if (n == 4 && p == 2) return (const void*)hvb_warp2_kernel<4, 2>;
    if (n == 5 && p == 2) return (const void*)hvb_warp2_kernel<5, 2>;
    if (n == 6 && p == 2) return (const void*)hvb_warp2_kernel<6, 2>;
    if (n == 7 && p == 2) return (const void*)hvb_warp2_kernel<7, 2>;
    if (n == 8 && p == 2) return (const void*)hvb_warp2_kernel<8, 2>;
    if (n == 2 && p == 4) return (const void*)hvb_warp2_kernel<2, 4>;
    if (n == 3 && p == 4) return (const void*)hvb_warp2_kernel<3, 4>;
    if (n == 4 && p == 4) return (const void*)hvb_warp2_kernel<4, 4>;
    if (n == 5 && p == 4) return (const void*)hvb_warp2_kernel<5, 4>;
    if (n == 6 && p == 4) return (const void*)hvb_warp2_kernel<6, 4>;
    if (n == 7 && p == 4) return (const void*)hvb_warp2_kernel<7, 4>;
    if (n == 8 && p == 4) return (const void*)hvb_warp2_kernel<8, 4>;
    if (n == 10 && p == 2) return (const void*)hvb_warp2_kernel<10, 2>;
    if (n == 12 && p == 2) return (const void*)hvb_warp2_kernel<12, 2>;
    if (n == 14 && p == 2) return (const void*)hvb_warp2_kernel<14, 2>;
    if (n == 16 && p == 2) return (const void*)hvb_warp2_kernel<16, 2>;
    if (n == 10 && p == 4) return (const void*)hvb_warp2_kernel<10, 4>;
    if (n == 12 && p == 4) return (const void*)hvb_warp2_kernel<12, 4>;
    if (n == 14 && p == 4) return (const void*)hvb_warp2_kernel<14, 4>;
    if (n == 16 && p == 4) return (const void*)hvb_warp2_kernel<16, 4>;
    if (n == 8 && p == 8) return (const void*)hvb_warp2_kernel<8, 8>;
    if (n == 10 && p == 8) return (const void*)hvb_warp2_kernel<10, 8>;
    if (n == 12 && p == 8) return (const void*)hvb_warp2_kernel<12, 8>;
    if (n == 14 && p == 8) return (const void*)hvb_warp2_kernel<14, 8>;
    return nullptr;
}
int hvb_warp2_group(int n) { return hvb_group2_for(n); }
