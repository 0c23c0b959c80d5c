// hvb_kernels.h -- host-visible kernel lookups (each .cu that instantiates kernels exports one).
#pragma once
#include <stddef.h>

const void* hvb_warp_kernel_ptr_g4(int pmax);
const void* hvb_warp_kernel_ptr_g8(int pmax);
const void* hvb_warp_kernel_ptr_g16(int pmax);
const void* hvb_warp_kernel_ptr_g32(int pmax);
int hvb_warp_threads();

const void* hvb_block_kernel_ptr(int nb);
int hvb_block_threads();
size_t hvb_block_smem_bytes(int nb, int n, int ncols, int region_a, int region_b);

const void* hvb_dump_kernel_ptr();
const void* hvb_dfma_probe_ptr();
