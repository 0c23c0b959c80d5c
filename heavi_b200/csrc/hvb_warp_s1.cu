// hvb_warp_s1.cu -- kernel A instantiations (generated list, see hvb_warp_shapes.h)
#include "hvb_warp.cuh"
#include "hvb_kernels.h"

const void* hvb_warp_kernel_ptr_s1(int n, int p) {
    if (n == 2 && p == 2) return (const void*)hvb_warp_kernel<2, 2>;
    if (n == 2 && p == 4) return (const void*)hvb_warp_kernel<2, 4>;
    if (n == 3 && p == 2) return (const void*)hvb_warp_kernel<3, 2>;
    if (n == 3 && p == 4) return (const void*)hvb_warp_kernel<3, 4>;
    if (n == 4 && p == 2) return (const void*)hvb_warp_kernel<4, 2>;
    if (n == 4 && p == 4) return (const void*)hvb_warp_kernel<4, 4>;
    if (n == 5 && p == 2) return (const void*)hvb_warp_kernel<5, 2>;
    if (n == 5 && p == 4) return (const void*)hvb_warp_kernel<5, 4>;
    return nullptr;
}
int hvb_warp_threads() { return HVB_WARP_THREADS; }
size_t hvb_warp_shared_words(int N, int nvc, int nnz_rows, int nfast, int P) {
    return hvb_warp_shared_words_dev(N, nvc, nnz_rows, nfast, P);
}
