// hvb_warp_s2.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_s2(int n, int p) {
    if (n == 6 && p == 2) return (const void*)hvb_warp_kernel<6, 2>;
    if (n == 6 && p == 4) return (const void*)hvb_warp_kernel<6, 4>;
    if (n == 7 && p == 2) return (const void*)hvb_warp_kernel<7, 2>;
    if (n == 7 && p == 4) return (const void*)hvb_warp_kernel<7, 4>;
    if (n == 8 && p == 2) return (const void*)hvb_warp_kernel<8, 2>;
    if (n == 8 && p == 4) return (const void*)hvb_warp_kernel<8, 4>;
    if (n == 8 && p == 8) return (const void*)hvb_warp_kernel<8, 8>;
    return nullptr;
}
