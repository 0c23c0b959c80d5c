// hvb_warp_b4.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_b4(int n, int p) {
    if (n == 32 && p == 2) return (const void*)hvb_warp_kernel<32, 2>;
    if (n == 32 && p == 4) return (const void*)hvb_warp_kernel<32, 4>;
    if (n == 32 && p == 8) return (const void*)hvb_warp_kernel<32, 8>;
    if (n == 32 && p == 16) return (const void*)hvb_warp_kernel<32, 16>;
    return nullptr;
}
