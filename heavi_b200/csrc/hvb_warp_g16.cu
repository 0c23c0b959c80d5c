// hvb_warp_g16.cu -- kernel A instantiations for groups of 16 lanes per system.
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_g16(int pmax) {
    switch (pmax) {
        case 2: return (const void*)hvb_warp_kernel<16, 2>;
        case 4: return (const void*)hvb_warp_kernel<16, 4>;
        case 8: return (const void*)hvb_warp_kernel<16, 8>;
        case 16: return (const void*)hvb_warp_kernel<16, 16>;
    }
    return nullptr;
}
