// hvb_warp_g8.cu -- kernel A instantiations for groups of 8 lanes per system.
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_g8(int pmax) {
    switch (pmax) {
        case 2: return (const void*)hvb_warp_kernel<8, 2>;
        case 4: return (const void*)hvb_warp_kernel<8, 4>;
        case 8: return (const void*)hvb_warp_kernel<8, 8>;
        case 16: return (const void*)hvb_warp_kernel<8, 16>;
    }
    return nullptr;
}
