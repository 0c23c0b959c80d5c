// cuda_shim.h -- TEST INFRASTRUCTURE: lets g++ compile the CUDA source that heavi_b200's kernel
// generator emits (hvb_codegen_source) so that its arithmetic can be checked on the CPU against the
// oracle without a GPU.  Nothing under heavi_b200/ uses this; the product runs the NVRTC-compiled kernel.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <cmath>

#define HVB_HOST_SIM 1
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __constant__ static const
#define __shared__ static

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct sim_dim3 { unsigned x, y, z; };
static sim_dim3 blockIdx{0, 0, 0}, threadIdx{0, 0, 0}, gridDim{1, 1, 1}, blockDim{1, 1, 1};

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline int atomicOr(int* p, int v) { const int o = *p; *p |= v; return o; }
using std::isfinite;
