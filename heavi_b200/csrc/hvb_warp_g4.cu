// hvb_warp_g4.cu -- kernel A instantiations for groups of 4 lanes per system.
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_g4(int pmax) {
    switch (pmax) {
        case 2: return (const void*)hvb_warp_kernel<4, 2>;
        case 4: return (const void*)hvb_warp_kernel<4, 4>;
        case 8: return (const void*)hvb_warp_kernel<4, 8>;
        case 16: return (const void*)hvb_warp_kernel<4, 16>;
    }
    return nullptr;
}
int hvb_warp_threads() { return HVB_WARP_THREADS; }
