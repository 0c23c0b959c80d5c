// hvb_jit.h -- NVRTC compile + module load/launch for generated kernels (hvb_jit.cpp).
#pragma once
#include <stddef.h>

#include <string>
#include <vector>

struct HvbJitKernel {
    void* module = nullptr;     // CUmodule
    void* func = nullptr;       // CUfunction
    int regs = 0, local_bytes = 0;
};

// source -> sm_100a cubin (needs libnvrtc, no GPU); log receives the compiler's messages
int hvb_jit_compile(const std::string& src, std::vector<char>* cubin, std::string* log, std::string* err);
// cubin -> module on the current device
int hvb_jit_load(const std::vector<char>& cubin, const char* name, HvbJitKernel* out, std::string* err);
void hvb_jit_unload(HvbJitKernel* k);
int hvb_jit_occupancy(const HvbJitKernel& k, int threads, size_t smem, int* blocks_per_sm, std::string* err);
int hvb_jit_set_max_dynamic_smem(const HvbJitKernel& k, size_t bytes, std::string* err);
int hvb_jit_launch(const HvbJitKernel& k, unsigned grid, unsigned threads, size_t smem, void* stream, void** args, std::string* err);
