// hvb_warp_b1.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_b1(int n, int p) {
    if (n == 20 && p == 2) return (const void*)hvb_warp_kernel<20, 2>;
    if (n == 20 && p == 4) return (const void*)hvb_warp_kernel<20, 4>;
    if (n == 20 && p == 8) return (const void*)hvb_warp_kernel<20, 8>;
    return nullptr;
}
