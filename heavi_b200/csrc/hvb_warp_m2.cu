// hvb_warp_m2.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_m2(int n, int p) {
    if (n == 14 && p == 2) return (const void*)hvb_warp_kernel<14, 2>;
    if (n == 14 && p == 4) return (const void*)hvb_warp_kernel<14, 4>;
    if (n == 14 && p == 8) return (const void*)hvb_warp_kernel<14, 8>;
    if (n == 16 && p == 2) return (const void*)hvb_warp_kernel<16, 2>;
    if (n == 16 && p == 4) return (const void*)hvb_warp_kernel<16, 4>;
    if (n == 16 && p == 8) return (const void*)hvb_warp_kernel<16, 8>;
    if (n == 16 && p == 16) return (const void*)hvb_warp_kernel<16, 16>;
    return nullptr;
}
