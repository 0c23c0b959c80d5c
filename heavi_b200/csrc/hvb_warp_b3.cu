// hvb_warp_b3.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_b3(int n, int p) {
    if (n == 28 && p == 2) return (const void*)hvb_warp_kernel<28, 2>;
    if (n == 28 && p == 4) return (const void*)hvb_warp_kernel<28, 4>;
    if (n == 28 && p == 8) return (const void*)hvb_warp_kernel<28, 8>;
    return nullptr;
}
