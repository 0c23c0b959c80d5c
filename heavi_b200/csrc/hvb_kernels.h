// hvb_kernels.h -- host-visible kernel lookups (each .cu that instantiates kernels exports one).
#pragma once
#include <stddef.h>

const void* hvb_warp_kernel_ptr_s1(int n, int p);
const void* hvb_warp_kernel_ptr_s2(int n, int p);
const void* hvb_warp_kernel_ptr_m1(int n, int p);
const void* hvb_warp_kernel_ptr_m2(int n, int p);
const void* hvb_warp_kernel_ptr_b1(int n, int p);
const void* hvb_warp_kernel_ptr_b2(int n, int p);
const void* hvb_warp_kernel_ptr_b3(int n, int p);
const void* hvb_warp_kernel_ptr_b4(int n, int p);
const void* hvb_warp2_kernel_ptr(int n, int p);
int hvb_warp2_group(int n);
int hvb_warp_threads();
size_t hvb_warp_shared_words(int N, int nvc, int nnz_rows, int nfast, int P);

const void* hvb_block_kernel_ptr(int nb, int minb);
int hvb_block_threads();
size_t hvb_block_smem_bytes(int nb, int n, int ncols, int region_a, int region_b);
size_t hvb_block_main_words(int nb, int n, int ncols, int region_a, int region_b);

const void* hvb_band_kernel_ptr(int b);
int hvb_band_threads();
int hvb_band_round_up(int b);
size_t hvb_band_words_per_thread(int b, int n, int pp, int n_dyn);
const void* hvb_bandf_kernel_ptr(int b, int twoport);
int hvb_bandf_ring_words(int b, int p);

const void* hvb_dump_kernel_ptr();
const void* hvb_stats_kernel_ptr();
int hvb_stats_threads();
const void* hvb_dfma_probe_ptr();
