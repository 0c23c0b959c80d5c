// hvb_jit.cpp -- run-time compilation of generated kernels: NVRTC (dlopen'ed, so that the library
// itself loads on machines without it) -> sm_100a cubin -> CUDA module through driver entry points
// obtained from the runtime (cudaGetDriverEntryPoint; no link-time dependency on libcuda).
#include "hvb_jit.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace {

// ---- NVRTC ---------------------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    std::string err;
};

Nvrtc* nvrtc() {
    static Nvrtc N;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> cands;
        const char* only = getenv("HVB_NVRTC_LIB");           // an explicit library is used exclusively
        if (only) cands.push_back(only);
        if (!only) cands.push_back("libnvrtc.so.12");
        if (!only) cands.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
        if (!only) cands.push_back("libnvrtc.so");
        // next to the CUDA runtime this process uses (pip layouts: .../nvidia/cuda_runtime/lib -> ../../cuda_nvrtc/lib)
        Dl_info di;
        if (!only && dladdr((void*)&cudaGetDeviceCount, &di) && di.dli_fname) {
            std::string p = di.dli_fname;
            const size_t s = p.rfind('/');
            if (s != std::string::npos) {
                const std::string dir = p.substr(0, s);
                cands.push_back(dir + "/libnvrtc.so.12");
                cands.push_back(dir + "/../../cuda_nvrtc/lib/libnvrtc.so.12");
            }
        }
        for (const std::string& c : cands) {
            N.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (N.h) break;
        }
        if (!N.h) { N.err = "libnvrtc.so.12 not found (set HVB_NVRTC_LIB)"; return; }
#define SYM(field, name) *(void**)(&N.field) = dlsym(N.h, name); if (!N.field) { N.err = std::string("missing NVRTC symbol ") + name; return; }
        SYM(CreateProgram, "nvrtcCreateProgram")
        SYM(CompileProgram, "nvrtcCompileProgram")
        SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        SYM(GetCUBIN, "nvrtcGetCUBIN")
        SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        SYM(GetProgramLog, "nvrtcGetProgramLog")
        SYM(DestroyProgram, "nvrtcDestroyProgram")
        SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
    });
    return &N;
}

// ---- driver entry points ---------------------------------------------------------------------------
typedef int CUresult;
typedef void* CUmodule;
typedef void* CUfunction;
struct Drv {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncGetAttribute)(int*, int, CUfunction) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, int, int) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    std::string err;
};

Drv* drv() {
    static Drv D;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [&](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
            if (e != cudaSuccess || !*fn) { D.err = std::string("driver entry point ") + name + " unavailable: " + cudaGetErrorString(e); return false; }
            return true;
        };
        if (!get("cuModuleLoadData", (void**)&D.ModuleLoadData)) return;
        if (!get("cuModuleUnload", (void**)&D.ModuleUnload)) return;
        if (!get("cuModuleGetFunction", (void**)&D.ModuleGetFunction)) return;
        if (!get("cuFuncGetAttribute", (void**)&D.FuncGetAttribute)) return;
        if (!get("cuFuncSetAttribute", (void**)&D.FuncSetAttribute)) return;
        if (!get("cuLaunchKernel", (void**)&D.LaunchKernel)) return;
        if (!get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&D.OccupancyMaxActiveBlocksPerMultiprocessor)) return;
        if (!get("cuGetErrorString", (void**)&D.GetErrorString)) return;
    });
    return &D;
}

std::string cu_err(CUresult r) {
    const char* s = nullptr;
    if (drv()->GetErrorString) drv()->GetErrorString(r, &s);
    return s ? s : "unknown driver error";
}

}  // namespace

int hvb_jit_compile(const std::string& src, std::vector<char>* cubin, std::string* log, std::string* err) {
    Nvrtc* N = nvrtc();
    if (!N->err.empty()) { if (err) *err = N->err; return -1; }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = N->CreateProgram(&prog, src.c_str(), "hvb_gen_kernel.cu", 0, nullptr, nullptr);
    if (r) { if (err) *err = std::string("nvrtcCreateProgram: ") + N->GetErrorString(r); return -1; }
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    std::string extra;
    if (const char* e = getenv("HVB_NVRTC_FLAGS")) extra = e;
    std::vector<std::string> extras;
    for (size_t i = 0; i < extra.size();) {
        const size_t j = extra.find(' ', i);
        const std::string tok = extra.substr(i, j == std::string::npos ? std::string::npos : j - i);
        if (!tok.empty()) extras.push_back(tok);
        if (j == std::string::npos) break;
        i = j + 1;
    }
    for (const std::string& x : extras) opts.push_back(x.c_str());
    r = N->CompileProgram(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    N->GetProgramLogSize(prog, &ls);
    std::string lg(ls, '\0');
    if (ls > 1) N->GetProgramLog(prog, &lg[0]);
    if (log) *log = lg;
    if (r) {
        if (err) *err = std::string("nvrtcCompileProgram: ") + N->GetErrorString(r) + "\n" + lg;
        N->DestroyProgram(&prog);
        return -1;
    }
    size_t cs = 0;
    r = N->GetCUBINSize(prog, &cs);
    if (r || cs == 0) { if (err) *err = "nvrtcGetCUBINSize failed"; N->DestroyProgram(&prog); return -1; }
    cubin->resize(cs);
    r = N->GetCUBIN(prog, cubin->data());
    N->DestroyProgram(&prog);
    if (r) { if (err) *err = "nvrtcGetCUBIN failed"; return -1; }
    return 0;
}

int hvb_jit_load(const std::vector<char>& cubin, const char* name, HvbJitKernel* out, std::string* err) {
    Drv* D = drv();
    if (!D->err.empty()) { if (err) *err = D->err; return -1; }
    cudaFree(0);                                       // make sure the runtime's primary context is current
    CUmodule mod = nullptr;
    CUresult r = D->ModuleLoadData(&mod, cubin.data());
    if (r) { if (err) *err = "cuModuleLoadData: " + cu_err(r); return -1; }
    CUfunction fn = nullptr;
    r = D->ModuleGetFunction(&fn, mod, name);
    if (r) { if (err) *err = "cuModuleGetFunction: " + cu_err(r); D->ModuleUnload(mod); return -1; }
    out->module = mod;
    out->func = fn;
    int v = 0;
    D->FuncGetAttribute(&v, 4 /*CU_FUNC_ATTRIBUTE_NUM_REGS*/, fn); out->regs = v;
    D->FuncGetAttribute(&v, 3 /*CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES*/, fn); out->local_bytes = v;
    return 0;
}

void hvb_jit_unload(HvbJitKernel* k) {
    if (k && k->module && drv()->ModuleUnload) drv()->ModuleUnload(k->module);
    if (k) { k->module = nullptr; k->func = nullptr; }
}

int hvb_jit_occupancy(const HvbJitKernel& k, int threads, size_t smem, int* blocks_per_sm, std::string* err) {
    CUresult r = drv()->OccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k.func, threads, smem);
    if (r) { if (err) *err = "cuOccupancyMaxActiveBlocksPerMultiprocessor: " + cu_err(r); return -1; }
    return 0;
}

int hvb_jit_set_max_dynamic_smem(const HvbJitKernel& k, size_t bytes, std::string* err) {
    CUresult r = drv()->FuncSetAttribute(k.func, 8 /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/, (int)bytes);
    if (r) { if (err) *err = "cuFuncSetAttribute(max dynamic shared memory): " + cu_err(r); return -1; }
    drv()->FuncSetAttribute(k.func, 9 /*CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT*/, 100);
    return 0;
}

int hvb_jit_launch(const HvbJitKernel& k, unsigned grid, unsigned threads, size_t smem, void* stream, void** args, std::string* err) {
    CUresult r = drv()->LaunchKernel(k.func, grid, 1, 1, threads, 1, 1, (unsigned)smem, stream, args, nullptr);
    if (r) { if (err) *err = "cuLaunchKernel: " + cu_err(r); return -1; }
    return 0;
}
