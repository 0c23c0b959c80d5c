// hvb_warp_g32.cu -- kernel A instantiations for groups of 32 lanes per system.
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_g32(int pmax) {
    switch (pmax) {
        case 2: return (const void*)hvb_warp_kernel<32, 2>;
        case 4: return (const void*)hvb_warp_kernel<32, 4>;
        case 8: return (const void*)hvb_warp_kernel<32, 8>;
        case 16: return (const void*)hvb_warp_kernel<32, 16>;
    }
    return nullptr;
}
