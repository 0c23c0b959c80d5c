// hvb_warp_b2.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_b2(int n, int p) {
    if (n == 24 && p == 2) return (const void*)hvb_warp_kernel<24, 2>;
    if (n == 24 && p == 4) return (const void*)hvb_warp_kernel<24, 4>;
    if (n == 24 && p == 8) return (const void*)hvb_warp_kernel<24, 8>;
    return nullptr;
}
