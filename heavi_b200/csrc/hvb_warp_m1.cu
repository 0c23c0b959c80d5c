// hvb_warp_m1.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_m1(int n, int p) {
    if (n == 10 && p == 2) return (const void*)hvb_warp_kernel<10, 2>;
    if (n == 10 && p == 4) return (const void*)hvb_warp_kernel<10, 4>;
    if (n == 10 && p == 8) return (const void*)hvb_warp_kernel<10, 8>;
    if (n == 12 && p == 2) return (const void*)hvb_warp_kernel<12, 2>;
    if (n == 12 && p == 4) return (const void*)hvb_warp_kernel<12, 4>;
    if (n == 12 && p == 8) return (const void*)hvb_warp_kernel<12, 8>;
    return nullptr;
}
