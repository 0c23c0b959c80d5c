// hvb_api.cu -- the C-ABI of libheavi_b200.so (include/heavi_b200.h): plan lifetime, kernel
// selection, the device-buffer entry (hvb_run_device) and the host-buffer pipeline (hvb_run_host).
#include <cuda_runtime.h>

#include <algorithm>
#include <functional>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include "../../include/heavi_b200.h"
#include "hvb_device.cuh"
#include "hvb_gen.h"
#include "hvb_jit.h"
#include "hvb_stamps.h"
#include "hvb_kernels.h"
#include "hvb_warp_shapes.h"

#include <deque>
#include <map>
#include <memory>
#include <mutex>

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(HVB_ERR_CUDA, "%s -> %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* hvb_last_error(void) { return g_err.c_str(); }
extern "C" const char* hvb_version(void) { return "heavi_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, need ? need : 16);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

// kernel D: one NVRTC-compiled kernel per (plan, parameter-reference layout)
struct GenKernel {
    HvbJitKernel k;
    DevBuf rows, itab, dtab;
    bool rolled = false, scratch_acc = false, two_sided = false;
    int occ_class = 0;                                   // 0: fewest spills, 1: one more resident CTA per SM (see ensure_gen_kernel)
    int ctas_per_sm = 0, n_cases = 0, threads = 128;
    size_t smem = 0;
    long long flops = 0;
    std::vector<int32_t> prefs;                        // the layout the value formulas were specialised on
    std::vector<double> consts;                        // ... and the constants baked into the code
};

struct hvb_plan {
    int device = 0;
    hvb_plan_desc d{};
    // device copies of the static plan arrays
    DevBuf const_vals, ops, coops, cell_ptr, cell_ent, row_ptr, row_ent, fop_code, fop_out, fop_arg, epi_por, epi_scale, port_rows, port_zpref, spl_t, spl_c, spl_trace, spl_fill, spl_range, band_slots, band_sched_ptr, band_sched_op, band_fslot, band_oslot, band_ent, band_act;
    // per-run parameter buffers (device) + pinned staging
    DevBuf consts, prefs;
    void* h_stage = nullptr; size_t h_stage_bytes = 0;
    void* h_in = nullptr; size_t h_in_bytes = 0;       // pinned staging for f / samples (pageable callers)
    int32_t* h_status = nullptr;                       // pinned status word
    cudaEvent_t ev_ready = nullptr, ev_tail = nullptr;
    // host pipeline buffers
    DevBuf f, samples, tables, S[2], status, work[2], dump, stats, inbuf, f32;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t upload_done = nullptr;
    // kernel choice
    int kclass = 0, G = 0, PMAX = 0, NB = 0, lanes = 0, rows_per_lane = 1, threads = 0, ctas_per_sm = 0, sm_count = 0, regs = 0;
    size_t smem = 0;
    int region_a = 0, region_b = 0;
    int band = -1;                                     // structural half bandwidth of the matrix block
    size_t work_words = 0;                             // kernel C: workspace words per thread
    int band_nslots = 0, band_min_row = 0, band_fused = 0;
    bool band_twoports_only = false;                   // every N-port block of the plan has two ports
    int band_value_slots = 0, band_nsm = 0;            // value slots the schedule needs / held in shared memory
    std::vector<int> rhs_minrow;                       // first row with an entry in each right-hand-side column
    long long band_rec_words = 0;
    const void* kernel = nullptr;
    int max_grid = 0;
    int64_t launches = 0;
    // kernel D (generated): host copy of the plan the generator reads, compiled variants, the active one
    std::vector<double> def_consts;                    // plans created from a stamp table: the run parameters it implies
    std::vector<int32_t> def_prefs;
    bool pending = false;                              // hvb_run_host_async queued, not yet waited for
    int gen_occ_matters = -1;                          // -1 unknown yet, 0 the generated kernel has one register budget, 1 two
    bool allow_two_sided = true;                       // chain kernels: cleared for good once a two-sided run flagged a point
    hvb_run_args pend_args{}; double* pend_out = nullptr;   // the queued call, should hvb_run_host_wait have to repeat it
    bool staged_valid = false;                         // h_in / inbuf hold the inputs of the last hvb_run_host call
    size_t staged_bytes = 0;
    int64_t staged_nF = 0, staged_nS = 0;
    std::vector<char> pushed;                          // consts + prefs as last uploaded by push_params
    bool pushed_valid = false;
    cudaStream_t pushed_stream = nullptr;
    GenPlan gen;
    bool gen_hub = false;                              // kernel E (N-port block + chains) instead of kernel D (one chain)
    std::vector<GenKernel*> gen_variants;
    GenKernel* gk = nullptr;
    // tunables read once, at plan creation (environment)
    int64_t env_chunk_mb = 0, env_min_slab = 0, env_stats_mb = 0;
    double env_tile_waves = 0.0;
    bool env_copy2d = true;
    int env_band_pfb = -1;
};

static cudaError_t upload(DevBuf& b, const void* src, size_t bytes) {
    cudaError_t e = b.ensure(bytes);
    if (e != cudaSuccess) return e;
    if (bytes) e = cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
    return e;
}

static const void* warp_kernel_ptr(int n, int p) {
    const void* k = nullptr;
    if (!k) k = hvb_warp_kernel_ptr_s1(n, p);
    if (!k) k = hvb_warp_kernel_ptr_s2(n, p);
    if (!k) k = hvb_warp_kernel_ptr_m1(n, p);
    if (!k) k = hvb_warp_kernel_ptr_m2(n, p);
    if (!k) k = hvb_warp_kernel_ptr_b1(n, p);
    if (!k) k = hvb_warp_kernel_ptr_b2(n, p);
    if (!k) k = hvb_warp_kernel_ptr_b3(n, p);
    if (!k) k = hvb_warp_kernel_ptr_b4(n, p);
    return k;
}

extern "C" int hvb_warp_shape(int32_t n, int32_t P, int32_t max_nport, int32_t* n_pad, int32_t* p_pad) {
    if (!n_pad || !p_pad) return fail(HVB_ERR_ARG, "null argument");
    *n_pad = 0; *p_pad = 0;
    const char* force = getenv("HVB_FORCE_BLOCK");
    if (force && force[0] == '1') return HVB_OK;
    const int need = std::max(std::max(n, max_nport), 2), needp = std::max(P, 1);
    double best = 0.0;
    for (int i = 0; i < HVB_N_WARP_SHAPES; ++i) {
        const int N = HVB_WARP_SHAPES[i][0], Pp = HVB_WARP_SHAPES[i][1];
        if (N < need || Pp < needp) continue;
        const int G = N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32;
        // issue-slot cost of one point: lanes occupied x elimination work per lane
        const double cost = (double)G * N * (0.5 * N + Pp + 6.0);
        if (*n_pad == 0 || cost < best) { best = cost; *n_pad = N; *p_pad = Pp; }
    }
    return HVB_OK;
}

// Kernel C's value schedule: evaluate each value op right before the first system row that sums
// one of its outputs, recycle the outputs' slots after the last such row (linear scan over rows, the
// lowest free slot first so that the busiest slots are the ones kept in shared memory).
static cudaError_t build_band_schedule(hvb_plan* pl, const hvb_plan_desc& d) {
    const int n = d.n, nvc = d.n_const_vals, ndyn = d.n_dyn;
    std::vector<int> first(ndyn, n), last(ndyn, -1);
    for (int r = 0; r < n; ++r)
        for (int e = d.row_ptr[r]; e < d.row_ptr[r + 1]; ++e) {
            const int vid = (d.row_ent[e] >> 1) & 0x7FFFF;
            if (vid >= nvc) { first[vid - nvc] = std::min(first[vid - nvc], r); last[vid - nvc] = std::max(last[vid - nvc], r); }
        }
    struct Op { int ref, out, cnt; };
    std::vector<std::vector<Op>> by_row(n);
    for (int o = 0; o < d.n_fast_ops; ++o) {
        const int v = d.fop_out[o];
        if (last[v] >= 0) by_row[first[v]].push_back({o, v, 1});
    }
    for (int o = 0; o < d.n_ops; ++o) {
        const int32_t* op = d.ops + (size_t)o * HVB_OP_WORDS;
        const int out = op[1] - nvc, cnt = op[0] == OP_TLINE ? 9 : 1;
        int fr = n;
        for (int e = 0; e < cnt; ++e) if (last[out + e] >= 0) fr = std::min(fr, first[out + e]);
        if (fr < n) by_row[fr].push_back({~o, out, cnt});
    }
    // two-port blocks ride in the same schedule, numbered after the simple ops: (Pn+1)^2 = 9 outputs each
    pl->band_twoports_only = true;
    for (int o = 0; o < d.n_coops; ++o) {
        const int32_t* op = d.coops + (size_t)o * HVB_COOP_WORDS;
        if (op[2] != 2) { pl->band_twoports_only = false; continue; }
        const int out = op[1] - nvc;
        int fr = n;
        for (int e = 0; e < 9; ++e) if (last[out + e] >= 0) fr = std::min(fr, first[out + e]);
        if (fr < n) by_row[fr].push_back({~(d.n_ops + o), out, 9});
    }
    std::vector<int> slot(ndyn, -1), sched_ptr(n + 1, 0), sched_op, fslot(std::max(d.n_fast_ops, 1), 0), oslot((size_t)std::max(d.n_ops + d.n_coops, 1) * 9, -1);
    std::vector<std::vector<int>> expire(n + 1);
    std::vector<int> free_slots;                           // min-heap
    int nslots = 0;
    for (int r = 0; r < n; ++r) {
        for (int v : expire[r]) { free_slots.push_back(slot[v]); std::push_heap(free_slots.begin(), free_slots.end(), std::greater<int>()); }
        for (const Op& op : by_row[r]) {
            sched_op.push_back(op.ref);
            for (int e = 0; e < op.cnt; ++e) {
                const int v = op.out + e;
                if (last[v] < 0) continue;
                int sl;
                if (!free_slots.empty()) { std::pop_heap(free_slots.begin(), free_slots.end(), std::greater<int>()); sl = free_slots.back(); free_slots.pop_back(); }
                else sl = nslots++;
                slot[v] = sl;
                expire[last[v] + 1].push_back(v);
                if (op.ref >= 0) fslot[op.ref] = sl; else oslot[(size_t)(~op.ref) * 9 + e] = sl;
            }
        }
        sched_ptr[r + 1] = (int)sched_op.size();
    }
    pl->rhs_minrow.assign((size_t)std::max(d.ncols - d.n, 1), n);
    for (int r = n - 1; r >= 0; --r)
        for (int e = d.row_ptr[r]; e < d.row_ptr[r + 1]; ++e) {
            const int col = (int)(((uint32_t)d.row_ent[e] >> 20) & 0x7FFu);
            if (col >= n) pl->rhs_minrow[col - n] = r;
        }
    std::vector<int32_t> ent((size_t)std::max(d.n_row_ent, 1), 0);
    for (int e = 0; e < d.n_row_ent; ++e) {
        const int32_t w = d.row_ent[e];
        const int vid = (w >> 1) & 0x7FFFF;
        if (vid < nvc) { ent[e] = w; continue; }
        const int nv = nvc + slot[vid - nvc];
        if (nv >= (1 << 19)) return cudaErrorInvalidValue;
        ent[e] = (int32_t)(((uint32_t)w & ~(0x7FFFFu << 1)) | ((uint32_t)nv << 1));
    }
    if (sched_op.empty()) sched_op.push_back(0);
    pl->band_value_slots = nslots;
    cudaError_t e = upload(pl->band_sched_ptr, sched_ptr.data(), sched_ptr.size() * 4);
    if (e == cudaSuccess) e = upload(pl->band_sched_op, sched_op.data(), sched_op.size() * 4);
    if (e == cudaSuccess) e = upload(pl->band_fslot, fslot.data(), fslot.size() * 4);
    if (e == cudaSuccess) e = upload(pl->band_oslot, oslot.data(), oslot.size() * 4);
    if (e == cudaSuccess) e = upload(pl->band_ent, ent.data(), ent.size() * 4);
    return e;
}

// kernel B's per-CTA workspaces; one set per pipeline slot because the two streams of hvb_run_host
// may have a kernel each in flight
static int ensure_block_work(hvb_plan* pl, int slot) {
    const hvb_plan_desc& d = pl->d;
    const size_t ldw = ((size_t)d.ncols + 3) & ~(size_t)3;
    const size_t need = (size_t)pl->max_grid * (d.n + 3) * ldw * sizeof(cplx);
    if (pl->work[slot].bytes >= need) return HVB_OK;
    CU(pl->work[slot].ensure(need));
    CU(cudaMemset(pl->work[slot].p, 0, pl->work[slot].bytes));   // padding rows / columns stay finite
    return HVB_OK;
}

static bool band_kernel_ok(const hvb_plan* pl) {
    const hvb_plan_desc& d = pl->d;
    const char* off = getenv("HVB_NO_BAND");
    if (off && off[0] == '1') return false;
    const char* force = getenv("HVB_FORCE_BLOCK");          // tests: dense block kernel for any shape
    if (force && force[0] == '1') return false;
    if (pl->band < 0 || d.kernel_hint == 1) return false;
    if (d.n_coops != 0) {
        // N-port blocks: only two-ports (evaluated per thread), and only the fused kernel knows how
        if (!pl->band_twoports_only || d.ncols - d.n != d.P) return false;
        const int Bc = hvb_band_round_up(pl->band);
        int smem_sm = 0;
        if (Bc < 1 || Bc > 4 || cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, pl->device) != cudaSuccess) return false;
        const int want_ctas = Bc == 1 ? 5 : Bc == 2 ? 4 : Bc == 3 ? 3 : 2;
        const size_t per_cta = (size_t)smem_sm / want_ctas - 1024;
        if ((size_t)(hvb_bandf_ring_words(Bc, d.P) + 1) * hvb_band_threads() * sizeof(cplx) > per_cta) return false;
        const char* nf = getenv("HVB_BAND_FUSED");
        if (nf && nf[0] == '0') return false;
    }
    const int B = hvb_band_round_up(pl->band);
    if (B == 0) return false;
    if (d.kernel_hint == 2) return B <= 4 && 2 * B + 1 <= d.n;   // the host measured / decided: any narrow band
    return 2 * (3 * B + 1) <= d.n;                       // a narrow band, not a small dense block
}

static bool gen_kernel_ok(const hvb_plan* pl) {
    const char* off = getenv("HVB_NO_GEN");
    if (off && off[0] == '1') return false;
    const char* nb = getenv("HVB_NO_BAND");                 // "dense kernels only" also rules out the chain kernel
    if (nb && nb[0] == '1') return false;
    const char* force = getenv("HVB_FORCE_BLOCK");
    if (force && force[0] == '1') return false;
    if (pl->d.kernel_hint == 1) return false;
    return hvb_gen_supported(pl->gen, nullptr) || hvb_gen_hub_candidate(pl->gen);
}

// ---- kernel D: generate + compile (cached per process by source text) + load ------------------------
static std::mutex g_cubin_mu;
static std::map<std::string, std::shared_ptr<std::vector<char>>> g_cubins;
static std::deque<std::string> g_cubin_order;            // insertion order: the cache holds the 256 most recent sources

// The generated code is specialised on the run's parameter layout (prefs) AND on the values of its constants (baked in
// as literals / constant-bank entries): a plan keeps one compiled variant per distinct (prefs, consts) it has seen.
//
// `points` = points of ONE launch when the caller launches the whole job at once (hvb_run_device), 0 for the host pipeline
// (its tiles are cut to one wave of the kernel).  The two-sided row loop exists in two register budgets: the compiler's
// own (168 registers for configs[4]: 384 threads per SM, fastest per thread; capped at 168 if the natural allocation is
// larger) and 128 registers (512 threads per SM, spills).  A launch of only a few waves is dominated by its last, partly
// filled wave, and there the higher occupancy wins -- measured on configs[4], 125 000 points (one GPU's share at
// N = 8): 0.445 ms -> 0.404 ms; 250 000: 0.784 -> 0.777; 1 000 000: 2.87 ms against 3.45.
static int ensure_gen_kernel(hvb_plan* pl, const int32_t* prefs, const double* consts, int64_t points = 0) {
    const size_t np = (size_t)pl->d.n_prefs * 2, nc = (size_t)pl->d.n_consts * 2;
    int cls = 0;
    if (pl->gen_occ_matters != 0 && pl->sm_count > 0) {
        const int64_t wave = (int64_t)3 * 128 * pl->sm_count;
        cls = (points > wave && points <= 6 * wave) ? 1 : 0;
    }
    auto same = [&](const GenKernel* v) {
        return v->occ_class == cls &&
               v->prefs.size() == np && (np == 0 || memcmp(v->prefs.data(), prefs, np * 4) == 0) &&
               v->consts.size() == nc && (nc == 0 || memcmp(v->consts.data(), consts, nc * 8) == 0);
    };
    if (pl->gk && same(pl->gk) && (pl->allow_two_sided || !pl->gk->two_sided)) return HVB_OK;
    for (GenKernel* v : pl->gen_variants)
        if (same(v) && (pl->allow_two_sided || !v->two_sided)) {
            pl->gk = v; pl->threads = v->threads; pl->smem = v->smem; pl->regs = v->k.regs; pl->ctas_per_sm = v->ctas_per_sm;
            pl->max_grid = v->ctas_per_sm * pl->sm_count;
            return HVB_OK;
        }
    if (pl->gen_variants.size() >= 64) {
        // a plan keeps at most 64 compiled variants; the oldest one that is not current goes (sample columns, not
        // constants, are the cheap way to change a value per run).  Its kernel may still be running on a caller's stream
        // (hvb_run_device), hence the device-wide wait before the module is unloaded.
        for (size_t i = 0; i < pl->gen_variants.size(); ++i) {
            GenKernel* v = pl->gen_variants[i];
            if (v == pl->gk) continue;
            CU(cudaDeviceSynchronize());
            v->rows.release(); v->itab.release(); v->dtab.release();
            hvb_jit_unload(&v->k);
            delete v;
            pl->gen_variants.erase(pl->gen_variants.begin() + (long)i);
            break;
        }
    }
    GenResult res;
    std::string err;
    pl->gen.run_consts.assign(consts, consts + nc);
    pl->gen.allow_two_sided = pl->allow_two_sided;
    pl->gen.occ_class = cls;
    const char* md = getenv("HVB_GEN_MODE");
    std::unique_ptr<GenKernel> gk(new GenKernel());
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int grc = pl->gen_hub ? hvb_gen_hub_source(pl->gen, prefs, hvb_gen_prelude(), &res, &err)
                                    : hvb_gen_source(pl->gen, prefs, md ? atoi(md) : 0, hvb_gen_prelude(), &res, &err);
        if (grc != 0) return fail(HVB_ERR_UNSUPPORTED, "kernel generator: %s", err.c_str());
        std::shared_ptr<std::vector<char>> cubin;
        {
            std::lock_guard<std::mutex> lk(g_cubin_mu);
            auto it = g_cubins.find(res.src);
            if (it != g_cubins.end()) cubin = it->second;
        }
        if (!cubin) {
            cubin = std::make_shared<std::vector<char>>();
            std::string log;
            if (hvb_jit_compile(res.src, cubin.get(), &log, &err) != 0) return fail(HVB_ERR_CUDA, "NVRTC: %.900s", err.c_str());
            std::lock_guard<std::mutex> lk(g_cubin_mu);
            if (g_cubins.emplace(res.src, cubin).second) g_cubin_order.push_back(res.src);
            while (g_cubin_order.size() > 256) { g_cubins.erase(g_cubin_order.front()); g_cubin_order.pop_front(); }
        }
        if (hvb_jit_load(*cubin, res.kernel_name.c_str(), &gk->k, &err) != 0) return fail(HVB_ERR_CUDA, "%s", err.c_str());
        // the compiler's own allocation is the fastest two-sided kernel as long as 384 threads stay resident per SM
        if (attempt == 0 && res.two_sided && cls == 0 && gk->k.regs > 168 && !getenv("HVB_GEN_MINB")) {
            hvb_jit_unload(&gk->k);
            pl->gen.occ_class = 2;
            continue;
        }
        break;
    }
    gk->rolled = res.rolled; gk->scratch_acc = res.scratch_acc; gk->n_cases = res.n_cases; gk->flops = res.flops_per_solve;
    gk->two_sided = res.two_sided;
    // only the two-sided row loop comes in two register budgets; for every other form the classes are one kernel
    pl->gen_occ_matters = res.two_sided ? 1 : 0;
    gk->occ_class = res.two_sided ? cls : 0;
    gk->prefs.assign(prefs, prefs + np);
    gk->consts.assign(consts, consts + nc);
    pl->threads = res.threads;
    gk->threads = res.threads; gk->smem = res.smem_bytes;
    if (res.smem_bytes > 48 * 1024) {
        std::string e2;
        if (hvb_jit_set_max_dynamic_smem(gk->k, res.smem_bytes, &e2) != 0) return fail(HVB_ERR_CUDA, "%s", e2.c_str());
    }
    if (res.rolled && !res.tables_in_source) {
        CU(upload(gk->rows, res.rows.data(), res.rows.size() * 4));
        CU(upload(gk->itab, res.itab.data(), res.itab.size() * 4));
        CU(upload(gk->dtab, res.dtab.data(), res.dtab.size() * 8));
    }
    int occ = 0;
    if (hvb_jit_occupancy(gk->k, pl->threads, res.smem_bytes, &occ, &err) != 0) return fail(HVB_ERR_CUDA, "%s", err.c_str());
    if (occ < 1) return fail(HVB_ERR_UNSUPPORTED, "generated kernel does not fit on an SM (%d registers)", gk->k.regs);
    gk->ctas_per_sm = occ;
    pl->regs = gk->k.regs;
    pl->ctas_per_sm = occ;
    pl->max_grid = occ * pl->sm_count;
    pl->smem = res.smem_bytes;
    pl->gk = gk.get();
    pl->gen_variants.push_back(gk.release());
    return HVB_OK;
}

static int select_kernel(hvb_plan* pl) {
    const hvb_plan_desc& d = pl->d;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, pl->device));
    pl->sm_count = prop.multiProcessorCount;
    int max_nport = 0;
    {
        // cooperative ops need one lane per N-port row
        std::vector<int32_t> co((size_t)d.n_coops * HVB_COOP_WORDS);
        if (d.n_coops) CU(cudaMemcpy(co.data(), pl->coops.p, co.size() * 4, cudaMemcpyDeviceToHost));
        for (int o = 0; o < d.n_coops; ++o) max_nport = std::max(max_nport, (int)co[(size_t)o * HVB_COOP_WORDS + 2]);
    }
    if (max_nport > HVB_MAX_NPORT) return fail(HVB_ERR_UNSUPPORTED, "N-port block with %d ports (max %d)", max_nport, HVB_MAX_NPORT);
    const int P = d.P;
    const int PP = d.ncols - d.n;
    pl->region_a = std::max(std::max(d.n_dyn, d.n * PP), 1);
    const void* wk = warp_kernel_ptr(d.n, PP);
    const char* force = getenv("HVB_FORCE_BLOCK");
    const size_t staged = (size_t)(d.n_const_vals + d.n_fast_ops) * 16 + (size_t)(2 * d.n + 1 + d.n_row_ent + 2 * d.n_fast_ops) * 4 + (size_t)d.P * d.P * 8 + 64;
    const bool band_first = d.kernel_hint == 2 && PP == d.P && band_kernel_ok(pl);
    if (gen_kernel_ok(pl)) {
        // kernel D: the netlist is a chain (tridiagonal in the plan's row order) -> a kernel generated for
        // exactly this netlist, one thread per point, compiled at the first run (hvb_gen.cpp, hvb_jit.cpp)
        pl->kclass = 3;
        pl->gen_hub = !hvb_gen_supported(pl->gen, nullptr);
        pl->NB = 1;
        pl->threads = pl->gen_hub ? 32 : 128;
        pl->smem = 0;
        pl->max_grid = pl->sm_count * 4;               // refined once the kernel exists
        return HVB_OK;
    }
    if (!band_first && wk && max_nport <= d.n && staged <= 24576 && !(force && force[0] == '1')) {
        const int N = d.n;
        const int G = N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32;
        pl->kclass = 0; pl->G = N; pl->PMAX = PP;
        pl->kernel = wk;
        pl->lanes = G; pl->rows_per_lane = 1;
        const char* r1 = getenv("HVB_WARP_R");
        // two rows per lane: the cooperative N-port op needs one lane per block row, so blocks up to the group size
        const void* wk2 = (max_nport <= hvb_warp2_group(N) && !(r1 && atoi(r1) == 1)) ? hvb_warp2_kernel_ptr(N, PP) : nullptr;
        if (wk2) { pl->kernel = wk2; pl->lanes = hvb_warp2_group(N); pl->rows_per_lane = 2; }
        pl->threads = hvb_warp_threads();
        pl->region_b = std::max(d.coop_scratch, N * (N + PP));
        const int sys_per_block = (pl->threads / 32) * (32 / pl->lanes);
        const size_t shared_words = hvb_warp_shared_words(N, d.n_const_vals, d.n_row_ent, d.n_fast_ops, d.P);
        pl->smem = (shared_words + (size_t)sys_per_block * (pl->region_a + pl->region_b)) * sizeof(cplx);
    } else if (band_kernel_ok(pl)) {
        // kernel C: thread per point, banded LU (hvb_band.cu); the workspace is sized per launch
        pl->kclass = 2;
        pl->NB = hvb_band_round_up(pl->band);
        pl->kernel = hvb_band_kernel_ptr(pl->NB);
        pl->threads = hvb_band_threads();
        pl->smem = 0;
        pl->work_words = hvb_band_words_per_thread(pl->NB, d.n, PP, d.n_dyn);
        const char* nf = getenv("HVB_BAND_FUSED");
        const void* fk = (PP == d.P && !(nf && nf[0] == '0')) ? hvb_bandf_kernel_ptr(pl->NB, d.n_coops > 0) : nullptr;
        if (fk) {
            // shared memory per thread: the right-hand-side ring + as many value slots as fit next to it
            const int want_ctas = pl->NB == 1 ? 5 : pl->NB == 2 ? 4 : pl->NB == 3 ? 3 : 2;
            const size_t per_cta = prop.sharedMemPerMultiprocessor / want_ctas - 1024;
            const size_t per_word = (size_t)pl->threads * sizeof(cplx);
            const int ring = hvb_bandf_ring_words(pl->NB, d.P);
            if ((size_t)(ring + 1) * per_word <= per_cta) {
                const int cap = (int)(per_cta / per_word) - ring;
                pl->kernel = fk;
                pl->band_fused = 1;
                pl->band_nsm = std::min(pl->band_value_slots, cap);
                pl->smem = (size_t)(ring + std::max(pl->band_nsm, 1)) * per_word;
                // right-hand-side columns that can be non-zero in rows <= k+B, and the record stream they give
                std::vector<int32_t> act((size_t)d.n, 0);
                long long words = 0;
                for (int k = 0; k < d.n; ++k) {
                    int a = 0;
                    for (int i = 0; i < d.P; ++i) if (pl->rhs_minrow[i] <= k + pl->NB) a = i + 1;
                    act[k] = a;
                    words += 2 * pl->NB + 1 + a;
                }
                pl->band_rec_words = words;
                CU(upload(pl->band_act, act.data(), act.size() * 4));
                pl->work_words = (size_t)words + (size_t)pl->band_nslots * d.P + (size_t)std::max(pl->band_value_slots - pl->band_nsm, 1);
            }
        }
    } else {
        pl->kclass = 1;
        const char* nbs = getenv("HVB_BLOCK_NB");
        pl->NB = (nbs && atoi(nbs) == 8) ? 8 : 16;
        pl->region_a = std::max(d.n_dyn, 1);             // solution rows live in the U12 buffer
        pl->region_b = std::max(d.coop_scratch, 1);
        if (d.n > 1023) return fail(HVB_ERR_UNSUPPORTED, "systems with more than 1023 unknowns are not supported (n=%d)", d.n);
        pl->threads = hvb_block_threads();
        pl->smem = hvb_block_smem_bytes(pl->NB, d.n, d.ncols, pl->region_a, pl->region_b);
        if (pl->smem > (size_t)prop.sharedMemPerBlockOptin && pl->NB == 16) {
            pl->NB = 8;
            pl->smem = hvb_block_smem_bytes(pl->NB, d.n, d.ncols, pl->region_a, pl->region_b);
        }
        {
            const char* mb = getenv("HVB_BLOCK_MINB");
            pl->kernel = hvb_block_kernel_ptr(pl->NB, (mb && atoi(mb) == 1) ? 1 : 2);
        }
    }
    if (!pl->kernel) return fail(HVB_ERR_UNSUPPORTED, "no kernel for n=%d P=%d", d.n, P);
    if (pl->smem > (size_t)prop.sharedMemPerBlockOptin)
        return fail(HVB_ERR_UNSUPPORTED, "system too large: needs %zu bytes of shared memory per CTA (max %zu)", pl->smem,
                    (size_t)prop.sharedMemPerBlockOptin);
    CU(cudaFuncSetAttribute(pl->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    CU(cudaFuncSetAttribute(pl->kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                            (pl->kclass == 2 && !pl->band_fused) ? (int)cudaSharedmemCarveoutMaxL1 : (int)cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, pl->kernel));
    pl->regs = fa.numRegs;
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pl->kernel, pl->threads, pl->smem));
    if (occ < 1) return fail(HVB_ERR_UNSUPPORTED, "kernel does not fit on an SM (smem %zu, regs %d)", pl->smem, pl->regs);
    pl->ctas_per_sm = occ;
    pl->max_grid = occ * pl->sm_count;
    if (pl->kclass == 1) {
        const char* cps = getenv("HVB_BLOCK_CTAS_PER_SM");
        if (cps && atoi(cps) > 0) pl->max_grid = std::min(pl->max_grid, atoi(cps) * pl->sm_count);
        { int rc = ensure_block_work(pl, 0); if (rc) return rc; }
    }
    if (pl->kclass == 2) {
        // bound the workspace: by default at most 24 GB per pipeline slot
        const char* wg = getenv("HVB_BAND_WORK_GB");
        const size_t budget = (size_t)(wg && atoi(wg) > 0 ? atoi(wg) : 24) << 30;
        const size_t per_cta = pl->work_words * sizeof(cplx) * pl->threads;
        pl->max_grid = (int)std::max<size_t>(std::min<size_t>((size_t)pl->max_grid, budget / per_cta), 1);
    }
    return HVB_OK;
}

extern "C" int hvb_plan_create(const hvb_plan_desc* desc, int device, hvb_plan** out) {
    if (!desc || !out) return fail(HVB_ERR_ARG, "null argument");
    if (desc->n < 1 || desc->ncols < desc->n || desc->P < 0) return fail(HVB_ERR_ARG, "bad sizes n=%d ncols=%d P=%d", desc->n, desc->ncols, desc->P);
    CU(cudaSetDevice(device));
    hvb_plan* pl = new hvb_plan();
    pl->device = device;
    pl->d = *desc;
    const hvb_plan_desc& d = *desc;
    const int ncell = d.n * d.ncols;
    const int nnz = d.cell_ptr[ncell];
    cudaError_t e = cudaSuccess;
    auto up = [&](DevBuf& b, const void* src, size_t bytes) { if (e == cudaSuccess) e = upload(b, src, bytes); };
    up(pl->const_vals, d.const_vals, (size_t)d.n_const_vals * 16);
    up(pl->ops, d.ops, (size_t)d.n_ops * HVB_OP_WORDS * 4);
    up(pl->coops, d.coops, (size_t)d.n_coops * HVB_COOP_WORDS * 4);
    up(pl->cell_ptr, d.cell_ptr, (size_t)(ncell + 1) * 4);
    up(pl->cell_ent, d.cell_ent, (size_t)nnz * 4);
    up(pl->row_ptr, d.row_ptr, (size_t)(d.n + 1) * 4);
    up(pl->row_ent, d.row_ent, (size_t)d.n_row_ent * 4);
    up(pl->fop_code, d.fop_code, (size_t)d.n_fast_ops * 4);
    up(pl->fop_out, d.fop_out, (size_t)d.n_fast_ops * 4);
    up(pl->fop_arg, d.fop_arg, (size_t)d.n_fast_ops * 8);
    up(pl->epi_por, d.epi_port_of_row, (size_t)d.n * 4);
    up(pl->epi_scale, d.epi_scale, (size_t)std::max(d.P * d.P, 1) * 8);
    up(pl->port_rows, d.port_rows, (size_t)d.P * 3 * 4);
    up(pl->port_zpref, d.port_zpref, (size_t)d.P * 4);
    up(pl->spl_t, d.spl_t, (size_t)d.n_knots * 8);
    up(pl->spl_c, d.spl_c, (size_t)d.n_coefs * 16);
    up(pl->spl_trace, d.spl_trace, (size_t)d.n_traces * 16);
    up(pl->spl_fill, d.spl_fill, (size_t)d.n_traces * 32);
    up(pl->spl_range, d.spl_range, (size_t)d.n_traces * 16);
    if (e == cudaSuccess) e = pl->consts.ensure((size_t)std::max(d.n_consts, 1) * 16);
    if (e == cudaSuccess) e = pl->prefs.ensure((size_t)std::max(d.n_prefs, 1) * 8);
    if (e == cudaSuccess) e = pl->status.ensure(16);
    pl->h_stage_bytes = (size_t)std::max(d.n_consts, 1) * 16 + (size_t)std::max(d.n_prefs, 1) * 8;
    if (e == cudaSuccess) e = cudaMallocHost(&pl->h_stage, pl->h_stage_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->streams[0], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->streams[1], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->upload_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_tail, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&pl->h_status, 16);
    if (e != cudaSuccess) {
        hvb_plan_destroy(pl);
        return fail(HVB_ERR_CUDA, "plan upload failed: %s", cudaGetErrorString(e));
    }
    {   // structural half bandwidth of the matrix block (right-hand-side columns excluded)
        int band = 0;
        for (int r = 0; r < d.n; ++r)
            for (int c = 0; c < d.n; ++c)
                if (d.cell_ptr[r * d.ncols + c + 1] > d.cell_ptr[r * d.ncols + c]) band = std::max(band, std::abs(r - c));
        pl->band = band;
    }
    {   // rows of the solution the S epilogue reads (port nodes): slot table for the fused band kernel
        std::vector<int32_t> slot((size_t)d.n, -1);
        int ns = 0, minrow = d.n;
        for (int i = 0; i < 3 * d.P; ++i) {
            const int r = d.port_rows[i];
            if (r >= 0 && r < d.n && slot[r] < 0) { slot[r] = ns++; minrow = std::min(minrow, r); }
        }
        pl->band_nslots = std::max(ns, 1);
        pl->band_min_row = ns ? minrow : d.n;
        if (upload(pl->band_slots, slot.data(), slot.size() * 4) != cudaSuccess ||
            build_band_schedule(pl, d) != cudaSuccess) {
            hvb_plan_destroy(pl);
            return fail(HVB_ERR_CUDA, "plan upload failed");
        }
    }
    pl->gen.from_desc(d);
    {   // environment tunables are read here, once, not on the per-call path
        const char* e;
        if ((e = getenv("HVB_CHUNK_MB")) && atoi(e) > 0) pl->env_chunk_mb = atoi(e);
        if ((e = getenv("HVB_MIN_SLAB")) && atoi(e) > 0) pl->env_min_slab = atoi(e);
        if ((e = getenv("HVB_STATS_MB")) && atoi(e) > 0) pl->env_stats_mb = atoi(e);
        if ((e = getenv("HVB_BAND_TILE_WAVES")) && atof(e) > 0) pl->env_tile_waves = atof(e);
        if ((e = getenv("HVB_COPY2D")) && e[0] == '0') pl->env_copy2d = false;
        if ((e = getenv("HVB_BAND_PFB"))) pl->env_band_pfb = atoi(e);
    }
    // the descriptor's host pointers are not kept
    pl->d.const_vals = nullptr; pl->d.ops = nullptr; pl->d.coops = nullptr; pl->d.cell_ptr = nullptr; pl->d.cell_ent = nullptr; pl->d.row_ptr = nullptr; pl->d.row_ent = nullptr; pl->d.fop_code = nullptr; pl->d.fop_out = nullptr; pl->d.fop_arg = nullptr; pl->d.epi_port_of_row = nullptr; pl->d.epi_scale = nullptr;
    pl->d.port_rows = nullptr; pl->d.port_zpref = nullptr; pl->d.spl_t = nullptr; pl->d.spl_c = nullptr; pl->d.spl_trace = nullptr;
    pl->d.spl_fill = nullptr; pl->d.spl_range = nullptr;
    if (d.ncols >= d.n + d.P && d.ncols > d.n) { // dump-only plans (no right-hand sides) never solve
        int rc = select_kernel(pl);
        if (rc != HVB_OK) { hvb_plan_destroy(pl); return rc; }
    }
    *out = pl;
    return HVB_OK;
}

extern "C" int hvb_plan_set_tunable(hvb_plan* pl, const char* name, double value) {
    if (!pl || !name) return fail(HVB_ERR_ARG, "null argument");
    const std::string n = name;
    if (n == "chunk_mb") pl->env_chunk_mb = (int64_t)value;
    else if (n == "min_slab") pl->env_min_slab = (int64_t)value;
    else if (n == "stats_mb") pl->env_stats_mb = (int64_t)value;
    else if (n == "band_tile_waves") pl->env_tile_waves = value;
    else if (n == "copy2d") pl->env_copy2d = value != 0.0;
    else if (n == "band_pfb") pl->env_band_pfb = (int)value;
    else if (n == "gen_two_sided") { pl->allow_two_sided = value != 0.0; pl->gk = nullptr; pl->gen_occ_matters = -1; }
    else return fail(HVB_ERR_ARG, "unknown tunable '%s'", name);
    return HVB_OK;
}

extern "C" void hvb_plan_destroy(hvb_plan* pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    DevBuf* all[] = {&pl->const_vals, &pl->ops, &pl->coops, &pl->cell_ptr, &pl->cell_ent, &pl->row_ptr, &pl->row_ent, &pl->fop_code, &pl->fop_out, &pl->fop_arg, &pl->epi_por, &pl->epi_scale, &pl->port_rows, &pl->port_zpref,
                     &pl->spl_t, &pl->spl_c, &pl->spl_trace, &pl->spl_fill, &pl->spl_range, &pl->band_slots, &pl->band_sched_ptr, &pl->band_sched_op, &pl->band_fslot, &pl->band_oslot, &pl->band_ent, &pl->band_act, &pl->consts, &pl->prefs,
                     &pl->f, &pl->samples, &pl->tables, &pl->S[0], &pl->S[1], &pl->status, &pl->work[0], &pl->work[1], &pl->f32, &pl->dump, &pl->stats, &pl->inbuf};
    for (DevBuf* b : all) b->release();
    for (GenKernel* v : pl->gen_variants) {
        v->rows.release(); v->itab.release(); v->dtab.release();
        hvb_jit_unload(&v->k);
        delete v;
    }
    pl->gen_variants.clear();
    if (pl->h_stage) cudaFreeHost(pl->h_stage);
    if (pl->h_in) cudaFreeHost(pl->h_in);
    if (pl->h_status) cudaFreeHost(pl->h_status);
    if (pl->ev_ready) cudaEventDestroy(pl->ev_ready);
    if (pl->ev_tail) cudaEventDestroy(pl->ev_tail);
    for (int i = 0; i < 2; ++i) if (pl->streams[i]) cudaStreamDestroy(pl->streams[i]);
    if (pl->upload_done) cudaEventDestroy(pl->upload_done);
    delete pl;
}

static HvbDev make_dev(const hvb_plan* pl) {
    HvbDev D{};
    const hvb_plan_desc& d = pl->d;
    D.n = d.n; D.ncols = d.ncols; D.P = d.P;
    D.PP = d.ncols - d.n; D.nnz_rows = d.n_row_ent;
    D.row_ptr = (const int*)pl->row_ptr.p;
    D.row_ent = (const int*)pl->row_ent.p;
    D.nfast = d.n_fast_ops; D.epi_simple = d.epi_simple;
    D.fop_code = (const int*)pl->fop_code.p;
    D.fop_out = (const int*)pl->fop_out.p;
    D.fop_arg = (const double*)pl->fop_arg.p;
    D.epi_port_of_row = (const int*)pl->epi_por.p;
    D.epi_scale = (const double*)pl->epi_scale.p;
    D.nvc = d.n_const_vals; D.ndyn = d.n_dyn; D.nops = d.n_ops; D.ncoops = d.n_coops; D.nR = d.n_samples_cols;
    D.coop_scratch = d.coop_scratch;
    D.region_a = pl->region_a; D.region_b = pl->region_b;
    D.block_main = pl->kclass == 1 ? (int)hvb_block_main_words(pl->NB, d.n, d.ncols, pl->region_a, pl->region_b) : 0;
    D.pv.consts = (const cplx*)pl->consts.p;
    D.pv.prefs = (const int2*)pl->prefs.p;
    D.const_vals = (const cplx*)pl->const_vals.p;
    D.ops = (const int*)pl->ops.p;
    D.coops = (const int*)pl->coops.p;
    D.cell_ptr = (const int*)pl->cell_ptr.p;
    D.cell_ent = (const int*)pl->cell_ent.p;
    D.port_rows = (const int*)pl->port_rows.p;
    D.port_zpref = (const int*)pl->port_zpref.p;
    D.pv.spl_t = (const double*)pl->spl_t.p;
    D.pv.spl_c = (const cplx*)pl->spl_c.p;
    D.pv.spl_trace = (const int4*)pl->spl_trace.p;
    D.pv.spl_fill = (const cplx*)pl->spl_fill.p;
    D.pv.spl_range = (const double2*)pl->spl_range.p;
    D.W = (cplx*)pl->work[0].p;
    D.band_slot_of_row = (const int*)pl->band_slots.p;
    D.band_nslots = pl->band_nslots; D.band_min_row = pl->band_min_row;
    {
        // records of small systems stay in the L2 (every thread rewrites the same words point after point):
        // prefetching them only costs issue slots
        const size_t footprint = (size_t)pl->band_rec_words * sizeof(cplx) * (size_t)pl->max_grid * pl->threads;
        D.band_pfb = pl->env_band_pfb >= 0 ? pl->env_band_pfb : (footprint > ((size_t)48 << 20) ? 2 : 0);
    }
    D.band_nsm = pl->band_nsm;
    D.band_sched_ptr = (const int*)pl->band_sched_ptr.p;
    D.band_sched_op = (const int*)pl->band_sched_op.p;
    D.band_fslot = (const int*)pl->band_fslot.p;
    D.band_oslot = (const int*)pl->band_oslot.p;
    D.band_ent = (const int*)pl->band_ent.p;
    D.band_act = (const int*)pl->band_act.p;
    D.band_rec_words = pl->band_rec_words;
    if (pl->gk) {
        D.gen_rows = (const int*)pl->gk->rows.p;
        D.gen_itab = (const int*)pl->gk->itab.p;
        D.gen_dtab = (const double*)pl->gk->dtab.p;
    }
    return D;
}

// stage consts + prefs through pinned memory so the copy is stream-ordered
static int push_params(hvb_plan* pl, const hvb_run_args* a, cudaStream_t st) {
    const hvb_plan_desc& d = pl->d;
    const size_t cb = (size_t)d.n_consts * 16, pb = (size_t)d.n_prefs * 8;
    if ((cb && !a->consts) || (pb && !a->prefs)) return fail(HVB_ERR_ARG, "consts/prefs missing");
    // unchanged since the last upload on this plan (the usual case for repeated sweeps): nothing to copy.
    // The device copy is only ever written by this function, stream-ordered before the kernels that read it.
    if (pl->pushed_valid && pl->pushed.size() == cb + pb && (cb == 0 || memcmp(pl->pushed.data(), a->consts, cb) == 0) &&
        (pb == 0 || memcmp(pl->pushed.data() + cb, a->prefs, pb) == 0) && pl->pushed_stream == st)
        return HVB_OK;
    // the previous run's copy out of the staging buffer must have finished
    CU(cudaEventSynchronize(pl->upload_done));
    pl->pushed.resize(cb + pb);
    if (cb) memcpy(pl->pushed.data(), a->consts, cb);
    if (pb) memcpy(pl->pushed.data() + cb, a->prefs, pb);
    pl->pushed_valid = true;
    pl->pushed_stream = st;
    char* hs = (char*)pl->h_stage;
    if (cb) memcpy(hs, a->consts, cb);
    if (pb) memcpy(hs + (size_t)std::max(d.n_consts, 1) * 16, a->prefs, pb);
    if (cb) CU(cudaMemcpyAsync(pl->consts.p, hs, cb, cudaMemcpyHostToDevice, st));
    if (pb) CU(cudaMemcpyAsync(pl->prefs.p, hs + (size_t)std::max(d.n_consts, 1) * 16, pb, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(pl->upload_done, st));
    return HVB_OK;
}

static int launch(hvb_plan* pl, HvbDev& D, cudaStream_t st, int slot = 0) {
    const long long total = D.nS * D.nF;
    if (total <= 0) return HVB_OK;
    long long blocks;
    if (pl->kclass == 0) {
        const int spb = (pl->threads / 32) * (32 / pl->lanes);
        blocks = (total + spb - 1) / spb;
    } else if (pl->kclass == 2 || pl->kclass == 3) {
        blocks = (total + pl->threads - 1) / pl->threads;
    } else {
        blocks = total;
    }
    const int grid = (int)std::min<long long>(blocks, pl->max_grid);
    if (pl->kclass == 1) {
        int rc = ensure_block_work(pl, slot);
        if (rc) return rc;
        D.W = (cplx*)pl->work[slot].p;
    } else if (pl->kclass == 2) {
        CU(pl->work[slot].ensure((size_t)grid * pl->threads * pl->work_words * sizeof(cplx)));
        D.W = (cplx*)pl->work[slot].p;
    }
    void* args[] = {&D};
    if (pl->kclass == 3) {
        if (!pl->gk) return fail(HVB_ERR_ARG, "generated kernel missing (internal)");
        std::string err;
        if (hvb_jit_launch(pl->gk->k, (unsigned)grid, (unsigned)pl->gk->threads, pl->gk->smem, st, args, &err) != 0) return fail(HVB_ERR_CUDA, "%s", err.c_str());
        pl->launches++;
        return HVB_OK;
    }
    CU(cudaLaunchKernel(pl->kernel, dim3(grid), dim3(pl->threads), args, pl->smem, st));
    pl->launches++;
    return HVB_OK;
}

// plans created from a stamp table carry their own consts / prefs: NULL in the run arguments means "those"
static hvb_run_args effective_args(const hvb_plan* pl, const hvb_run_args* a) {
    hvb_run_args b = *a;
    if (!b.consts && !pl->def_consts.empty()) b.consts = pl->def_consts.data();
    if (!b.prefs && !pl->def_prefs.empty()) b.prefs = pl->def_prefs.data();
    return b;
}

static int check_args(const hvb_plan* pl, const hvb_run_args* a) {
    if (!pl || !a) return fail(HVB_ERR_ARG, "null argument");
    if (a->nF < 0 || a->nS < 1) return fail(HVB_ERR_ARG, "bad nF=%lld nS=%lld", (long long)a->nF, (long long)a->nS);
    if (pl->d.n_samples_cols > 0 && !a->samples) return fail(HVB_ERR_ARG, "plan has %d sample columns but samples is NULL", pl->d.n_samples_cols);
    if (a->n_tables > 0 && !a->tables) return fail(HVB_ERR_ARG, "tables missing");
    return HVB_OK;
}

extern "C" int hvb_run_device(hvb_plan* pl, const hvb_run_args* a_in, double* d_S, int64_t ldS, int32_t* d_status, void* stream) {
    if (!pl || !a_in) return fail(HVB_ERR_ARG, "null argument");
    const hvb_run_args a_eff = effective_args(pl, a_in);
    const hvb_run_args* a = &a_eff;
    int rc = check_args(pl, a);
    if (rc) return rc;
    if (pl->d.ncols <= pl->d.n || (!pl->kernel && pl->kclass != 3)) return fail(HVB_ERR_ARG, "plan has no right-hand sides (dump plan)");
    CU(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->kclass == 3 && (rc = ensure_gen_kernel(pl, a->prefs, a->consts, a->nF * a->nS)) != HVB_OK) return rc;
    rc = push_params(pl, a, st);
    if (rc) return rc;
    HvbDev D = make_dev(pl);
    D.nF = a->nF; D.nS = a->nS; D.ldS = ldS; D.pv.table_ld = a->nF;
    D.f = a->f; D.samples = a->samples; D.pv.tables = (const cplx*)a->tables;
    D.S = (cplx*)d_S; D.status = d_status;
    return launch(pl, D, st);
}

static int run_host_impl(hvb_plan* pl, const hvb_run_args* a, double* S_out, int32_t* status_out, bool wait);

static void drain_after_error(hvb_plan* pl) {
    // an error path may leave kernels and copies of earlier tiles in flight: they read the pinned staging
    // block and write the caller's S_out, so wait for both pipeline streams before handing control back
    const std::string keep = g_err;
    cudaSetDevice(pl->device);
    cudaStreamSynchronize(pl->streams[0]);
    cudaStreamSynchronize(pl->streams[1]);
    cudaGetLastError();
    g_err = keep;
}

// Non-blocking form: everything is queued on the plan's streams and the call returns; hvb_run_host_wait blocks
// until S_out (and f32_out) are complete.  Several plans -- e.g. the three filters of BASELINE configs[1] -- can
// have a run in flight each, so their copies share the PCIe link back to back instead of each call's fixed
// costs (host marshalling, first kernel, last copy) adding up.  Inputs are staged during the call; S_out,
// f32_out and a page-locked `tables` array must stay alive until the wait.
// The two-sided chain kernel flags a point it cannot serve safely (HVB_STATUS_RERUN_BIT, hvb_gen.cpp).  The plan then
// goes back to the one-sided kernel for good and the call is repeated: the bit never reaches a caller of the host entries.
static bool wants_rerun(hvb_plan* pl, int32_t st) {
    if (!(st & (int32_t)HVB_STATUS_RERUN_BIT)) return false;
    pl->allow_two_sided = false;
    pl->gk = nullptr;
    return true;
}

extern "C" int hvb_run_host_async(hvb_plan* pl, const hvb_run_args* a, double* S_out) {
    if (pl && pl->pending) return fail(HVB_ERR_ARG, "a run is already in flight on this plan (hvb_run_host_wait first)");
    const int rc = run_host_impl(pl, a, S_out, nullptr, false);
    if (rc != HVB_OK && pl && pl->streams[0]) drain_after_error(pl);
    if (rc == HVB_OK && pl) { pl->pending = true; pl->pend_args = effective_args(pl, a); pl->pend_out = S_out; }
    return rc;
}

extern "C" int hvb_run_host_wait(hvb_plan* pl, int32_t* status_out) {
    if (!pl) return fail(HVB_ERR_ARG, "null argument");
    if (!pl->pending) { if (status_out) *status_out = 0; return HVB_OK; }
    pl->pending = false;
    CU(cudaSetDevice(pl->device));
    CU(cudaStreamSynchronize(pl->streams[0]));
    int32_t st = *pl->h_status;
    if (wants_rerun(pl, st)) {
        // repeat from the staged copy of the inputs (the caller's f / samples / consts need not be alive any more)
        const hvb_plan_desc& d = pl->d;
        auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
        hvb_run_args b = pl->pend_args;
        char* h = (char*)pl->h_in;
        const size_t off_consts = 16, off_prefs = off_consts + up16((size_t)d.n_consts * 16), off_smp = off_prefs + up16((size_t)d.n_prefs * 8);
        const size_t off_f = off_smp + up16(d.n_samples_cols ? (size_t)b.nS * d.n_samples_cols * 8 : 0);
        b.consts = (const double*)(h + off_consts); b.prefs = (const int32_t*)(h + off_prefs);
        b.samples = d.n_samples_cols ? (const double*)(h + off_smp) : nullptr;
        b.f = (const double*)(h + off_f);
        const int rc = run_host_impl(pl, &b, pl->pend_out, &st, true);
        if (rc != HVB_OK) { drain_after_error(pl); return rc; }
    }
    if (status_out) *status_out = st & ~(int32_t)HVB_STATUS_RERUN_BIT;
    return HVB_OK;
}

extern "C" int hvb_run_host(hvb_plan* pl, const hvb_run_args* a, double* S_out, int32_t* status_out) {
    if (pl && pl->pending) return fail(HVB_ERR_ARG, "a run is already in flight on this plan (hvb_run_host_wait first)");
    int32_t st = 0;
    int rc = run_host_impl(pl, a, S_out, &st, true);
    if (rc == HVB_OK && wants_rerun(pl, st)) rc = run_host_impl(pl, a, S_out, &st, true);
    if (rc != HVB_OK && pl && pl->streams[0]) drain_after_error(pl);
    if (status_out) *status_out = st & ~(int32_t)HVB_STATUS_RERUN_BIT;
    return rc;
}

static int run_host_impl(hvb_plan* pl, const hvb_run_args* a_in, double* S_out, int32_t* status_out, bool wait) {
    if (!pl || !a_in) return fail(HVB_ERR_ARG, "null argument");
    const hvb_run_args a_eff = effective_args(pl, a_in);
    const hvb_run_args* a = &a_eff;
    int rc = check_args(pl, a);
    if (rc) return rc;
    if (!S_out) return fail(HVB_ERR_ARG, "S_out is NULL");
    const hvb_plan_desc& d = pl->d;
    if (d.ncols <= d.n || (!pl->kernel && pl->kclass != 3)) return fail(HVB_ERR_ARG, "plan has no right-hand sides (dump plan)");
    CU(cudaSetDevice(pl->device));
    const int64_t nF = a->nF, nS = a->nS;
    const int64_t PP = (int64_t)d.P * d.P;
    if (nF == 0 || PP == 0) { if (status_out) *status_out = 0; return HVB_OK; }
    if ((d.n_consts && !a->consts) || (d.n_prefs && !a->prefs)) return fail(HVB_ERR_ARG, "consts/prefs missing");
    if (pl->kclass == 3 && (rc = ensure_gen_kernel(pl, a->prefs, a->consts)) != HVB_OK) return rc;
    cudaStream_t s0 = pl->streams[0], s1 = pl->streams[1];
    if (a->out_ld != 0 && a->out_ld < nF) return fail(HVB_ERR_ARG, "out_ld=%lld is smaller than nF=%lld", (long long)a->out_ld, (long long)nF);
    const int64_t ldo = a->out_ld ? a->out_ld : nF;       // frequency pitch of the caller's array
    static const bool trace = getenv("HVB_TRACE") != nullptr;
    timespec tsp[8]; int ntp = 0;
    auto mark = [&]() { if (trace && ntp < 8) clock_gettime(CLOCK_MONOTONIC, &tsp[ntp++]); };
    mark();

    // ---- tiles: whole samples when several fit in a chunk, else frequency slabs of one sample --------
    // Tile size: large enough for the copy to run at the full PCIe rate (a few MB), small enough that
    // the first copy starts early and the exposed head (first kernel) / tail stay short; tiles are
    // equal-sized so no short tile trails the pipeline.
    int64_t chunk_bytes = (pl->env_chunk_mb > 0 ? pl->env_chunk_mb : 16) << 20;
    // the thread-per-point kernels: a tile should fill the grid at least once
    if ((pl->kclass == 2 || pl->kclass == 3) && pl->env_chunk_mb <= 0) {
        const double waves = pl->env_tile_waves > 0 ? pl->env_tile_waves : 1.0;
        chunk_bytes = std::max<int64_t>(chunk_bytes, (int64_t)(waves * pl->max_grid * pl->threads) * PP * 16);
    }
    // points per slab at least (a slab leaves as P*P segments of 16*slab bytes).  Measured: a 100 k-point
    // two-port sweep is fastest as two slabs, a 1 M-point eight-port sweep with ~30 (shorter exposed tail)
    const int64_t min_slab = pl->env_min_slab > 0 ? pl->env_min_slab : (nS * nF < 200000 ? 49152 : 32768);
    const int64_t total_bytes = nS * nF * PP * 16;
    int64_t ntiles = std::max<int64_t>((total_bytes + chunk_bytes - 1) / chunk_bytes, 1);
    // small jobs: still cut them so that the copy of one tile overlaps the kernel of the next
    if (ntiles < 4 && pl->kclass != 2 && pl->kclass != 3) ntiles = std::max<int64_t>(ntiles, std::min<int64_t>(4, nS * nF / min_slab));
    int64_t Sc = nS, Fc = nF;
    if (nS >= ntiles) {
        Sc = (nS + ntiles - 1) / ntiles;                  // whole samples per tile
    } else {
        Sc = 1;                                           // frequency slabs of one sample
        int64_t slabs = (ntiles + nS - 1) / nS;
        slabs = std::max<int64_t>(std::min<int64_t>(slabs, nF / min_slab), 1);
        Fc = (nF + slabs - 1) / slabs;
    }
    for (int b = 0; b < 2; ++b) CU(pl->S[b].ensure((size_t)Sc * PP * Fc * 16));

    // ---- inputs: ONE pinned staging block [status | consts | prefs | samples | f] -------------------
    // (f and the sample matrix usually come from pageable numpy memory, where a direct cudaMemcpyAsync
    // degrades to a synchronous staged copy, and every extra API call costs microseconds.)  A single
    // sweep cut into frequency slabs uploads f slab by slab, each on its tile's stream, so the first
    // kernel starts after the first slab instead of after the whole axis.
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t off_consts = 16;
    const size_t off_prefs = off_consts + up16((size_t)d.n_consts * 16);
    const size_t off_smp = off_prefs + up16((size_t)d.n_prefs * 8);
    const size_t off_f = off_smp + up16(d.n_samples_cols ? (size_t)nS * d.n_samples_cols * 8 : 0);
    const size_t in_bytes = off_f + up16((size_t)nF * 8);
    const bool slab_upload = nS == 1 && Fc < nF;
    if (pl->h_in_bytes < in_bytes) {
        CU(cudaStreamSynchronize(s0));
        if (pl->h_in) cudaFreeHost(pl->h_in);
        pl->h_in = nullptr; pl->h_in_bytes = 0;
        CU(cudaMallocHost(&pl->h_in, in_bytes + in_bytes / 4 + 4096));
        pl->h_in_bytes = in_bytes + in_bytes / 4 + 4096;
    }
    CU(pl->inbuf.ensure(in_bytes));
    if (a->f32_out) CU(pl->f32.ensure((size_t)nF * 4));
    char* h = (char*)pl->h_in;                            // free: the previous call waited for all its copies
    const int64_t f_first = slab_upload ? Fc : nF;
    // Repeated sweeps (an optimiser's inner loop, a GUI, the benchmark) usually come with the same frequencies,
    // samples and parameters: when the whole input block equals what the previous call staged -- and the device
    // copy is therefore still current -- only the status word is reset and nothing is uploaded.
    const bool same_inputs = pl->staged_valid && pl->staged_bytes == in_bytes && pl->staged_nF == nF && pl->staged_nS == nS &&
        (!d.n_consts || memcmp(h + off_consts, a->consts, (size_t)d.n_consts * 16) == 0) &&
        (!d.n_prefs || memcmp(h + off_prefs, a->prefs, (size_t)d.n_prefs * 8) == 0) &&
        (!d.n_samples_cols || memcmp(h + off_smp, a->samples, (size_t)nS * d.n_samples_cols * 8) == 0) &&
        memcmp(h + off_f, a->f, (size_t)nF * 8) == 0;
    if (same_inputs) {
        CU(cudaMemsetAsync(pl->inbuf.p, 0, 16, s0));
    } else {
        pl->staged_valid = false;
        memset(h, 0, 16);
        if (d.n_consts) memcpy(h + off_consts, a->consts, (size_t)d.n_consts * 16);
        if (d.n_prefs) memcpy(h + off_prefs, a->prefs, (size_t)d.n_prefs * 8);
        if (d.n_samples_cols) memcpy(h + off_smp, a->samples, (size_t)nS * d.n_samples_cols * 8);
        memcpy(h + off_f, a->f, (size_t)f_first * 8);
        CU(cudaMemcpyAsync(pl->inbuf.p, h, off_f + (size_t)f_first * 8, cudaMemcpyHostToDevice, s0));
    }
    if (a->n_tables) {
        CU(pl->tables.ensure((size_t)a->n_tables * nF * 16));
        CU(cudaMemcpyAsync(pl->tables.p, a->tables, (size_t)a->n_tables * nF * 16, cudaMemcpyHostToDevice, s0));
    }
    CU(cudaEventRecord(pl->ev_ready, s0));
    CU(cudaStreamWaitEvent(s1, pl->ev_ready, 0));
    char* dbase = (char*)pl->inbuf.p;
    mark();

    const bool copy2d = pl->env_copy2d;                   // default; HVB_COPY2D=0 = one plain copy per row (measured slower)
    int tile = 0;
    for (int64_t sb = 0; sb < nS; sb += Sc) {
        const int64_t ns = std::min(Sc, nS - sb);
        for (int64_t fb = 0; fb < nF; fb += Fc, ++tile) {
            const int64_t nf = std::min(Fc, nF - fb);
            cudaStream_t st = pl->streams[tile & 1];
            if (slab_upload && fb > 0 && !same_inputs) {
                memcpy(h + off_f + (size_t)fb * 8, a->f + fb, (size_t)nf * 8);
                CU(cudaMemcpyAsync(dbase + off_f + (size_t)fb * 8, h + off_f + (size_t)fb * 8, (size_t)nf * 8,
                                   cudaMemcpyHostToDevice, st));
            }
            HvbDev D = make_dev(pl);
            D.nF = nf; D.nS = ns; D.ldS = nf; D.pv.table_ld = nF;
            D.pv.consts = (const cplx*)(dbase + off_consts);
            D.pv.prefs = (const int2*)(dbase + off_prefs);
            D.f = (const double*)(dbase + off_f) + fb;
            D.samples = d.n_samples_cols ? (const double*)(dbase + off_smp) + sb * d.n_samples_cols : nullptr;
            D.pv.tables = a->n_tables ? (const cplx*)pl->tables.p + fb : nullptr;
            D.S = (cplx*)pl->S[tile & 1].p;
            D.status = (int*)pl->inbuf.p;
            // the points of sample 0 also round their frequency to float32 (`Sparameters.f`, network.py:407,421)
            D.f32 = (a->f32_out && sb == 0) ? (float*)pl->f32.p + fb : nullptr;
            rc = launch(pl, D, st, tile & 1);
            if (rc) return rc;
            // rows of nf elements, one per (sample, j, i); destination pitch is the caller's (default: the full nF)
            char* dst = (char*)S_out + ((size_t)sb * PP * ldo + (size_t)fb) * 16;
            // contiguous destination -> one plain copy; otherwise one strided 2-D copy
            const int64_t rows = ns * PP;
            if (nf == ldo) {
                CU(cudaMemcpyAsync(dst, D.S, (size_t)rows * nf * 16, cudaMemcpyDeviceToHost, st));
            } else if (rows <= 64 && !copy2d) {
                for (int64_t rr = 0; rr < rows; ++rr)
                    CU(cudaMemcpyAsync(dst + (size_t)rr * ldo * 16, (const char*)D.S + (size_t)rr * nf * 16, (size_t)nf * 16,
                                       cudaMemcpyDeviceToHost, st));
            } else {
                CU(cudaMemcpy2DAsync(dst, (size_t)ldo * 16, D.S, (size_t)nf * 16, (size_t)nf * 16, (size_t)rows,
                                     cudaMemcpyDeviceToHost, st));
            }
        }
    }
    // stream 0 waits for stream 1's tail, fetches the status word, and is the only stream the host waits on
    CU(cudaEventRecord(pl->ev_tail, s1));
    CU(cudaStreamWaitEvent(s0, pl->ev_tail, 0));
    CU(cudaMemcpyAsync(pl->h_status, pl->inbuf.p, 4, cudaMemcpyDeviceToHost, s0));
    if (a->f32_out) CU(cudaMemcpyAsync(a->f32_out, pl->f32.p, (size_t)nF * 4, cudaMemcpyDeviceToHost, s0));   // one copy for the whole axis
    mark();
    mark();
    pl->staged_valid = true; pl->staged_bytes = in_bytes; pl->staged_nF = nF; pl->staged_nS = nS;   // every upload of this call is queued
    if (!wait) return HVB_OK;                              // hvb_run_host_async: the caller collects with hvb_run_host_wait
    CU(cudaStreamSynchronize(s0));
    mark();
    if (trace) {
        auto us = [&](int i) { return (tsp[i].tv_sec - tsp[0].tv_sec) * 1e6 + (tsp[i].tv_nsec - tsp[0].tv_nsec) * 1e-3; };
        fprintf(stderr, "hvb_run_host: staged+H2D queued %.1f us, tiles queued %.1f, synced %.1f (tiles=%d)\n", us(1), us(2), us(4), tile);
    }
    if (status_out) *status_out = *pl->h_status;
    return HVB_OK;
}

extern "C" int hvb_run_host_multi(hvb_plan* const* plans, int32_t nplans, const hvb_run_args* a, double* S_out, int32_t* status_out) {
    if (!plans || nplans < 1 || !a) return fail(HVB_ERR_ARG, "null argument");
    for (int k = 0; k < nplans; ++k) {
        if (!plans[k]) return fail(HVB_ERR_ARG, "null plan");
        if (plans[k]->d.n != plans[0]->d.n || plans[k]->d.P != plans[0]->d.P || plans[k]->d.n_prefs != plans[0]->d.n_prefs)
            return fail(HVB_ERR_ARG, "plans are not replicas of one netlist");
    }
    if (nplans == 1) return hvb_run_host(plans[0], a, S_out, status_out);
    if (a->out_ld != 0) return fail(HVB_ERR_ARG, "out_ld must be 0 for the multi-device entry");
    const int64_t nF = a->nF, nS = a->nS;
    if (nF < 0 || nS < 1) return fail(HVB_ERR_ARG, "bad nF=%lld nS=%lld", (long long)nF, (long long)nS);
    const int64_t PP = (int64_t)plans[0]->d.P * plans[0]->d.P;
    const bool by_sample = nS >= nplans;
    const int64_t total = by_sample ? nS : nF;
    std::vector<int> rcs((size_t)nplans, HVB_OK);
    std::vector<int32_t> sts((size_t)nplans, 0);
    std::vector<std::string> errs((size_t)nplans);
    std::vector<std::thread> threads;
    auto block = [&](int k, int64_t* lo, int64_t* hi) {          // same contiguous near-equal split as sharding.split_range
        const int64_t base = total / nplans, extra = total % nplans;
        *lo = k * base + std::min<int64_t>(k, extra);
        *hi = *lo + base + (k < extra ? 1 : 0);
    };
    for (int k = 0; k < nplans; ++k) {
        threads.emplace_back([&, k]() {
            int64_t lo, hi;
            block(k, &lo, &hi);
            if (hi <= lo) return;
            hvb_run_args b = *a;
            double* out = S_out;
            if (by_sample) {
                b.nS = hi - lo;
                if (a->samples) b.samples = a->samples + lo * plans[k]->d.n_samples_cols;
                out = S_out + (size_t)lo * PP * nF * 2;
                if (lo != 0) b.f32_out = nullptr;               // the float32 axis comes from the block that holds sample 0
            } else {
                b.nF = hi - lo;
                b.f = a->f + lo;
                if (a->tables) b.tables = nullptr;               // handled below: tables are per-frequency rows of pitch nF
                out = S_out + (size_t)lo * 2;
                b.out_ld = nF;
                if (a->f32_out) b.f32_out = a->f32_out + lo;
            }
            std::vector<double> tab;
            if (!by_sample && a->n_tables > 0) {                 // this slab's columns of every table, made contiguous
                tab.resize((size_t)a->n_tables * (hi - lo) * 2);
                for (int t = 0; t < a->n_tables; ++t)
                    memcpy(&tab[(size_t)t * (hi - lo) * 2], a->tables + ((size_t)t * nF + lo) * 2, (size_t)(hi - lo) * 16);
                b.tables = tab.data();
            }
            rcs[k] = hvb_run_host(plans[k], &b, out, &sts[k]);
            if (rcs[k] != HVB_OK) errs[k] = g_err;               // g_err is thread-local
        });
    }
    for (std::thread& t : threads) t.join();
    int32_t st = 0;
    for (int k = 0; k < nplans; ++k) {
        if (rcs[k] != HVB_OK) return fail(rcs[k], "device %d: %s", plans[k]->device, errs[k].c_str());
        st |= sts[k];
    }
    if (status_out) *status_out = st;
    return HVB_OK;
}

static int run_stats_host_impl(hvb_plan* pl, const hvb_run_args* a_in, double* stats_out, int32_t* status_out);

extern "C" int hvb_run_stats_host(hvb_plan* pl, const hvb_run_args* a, double* stats_out, int32_t* status_out) {
    int32_t st = 0;
    int rc = run_stats_host_impl(pl, a, stats_out, &st);
    if (rc == HVB_OK && wants_rerun(pl, st)) rc = run_stats_host_impl(pl, a, stats_out, &st);
    if (status_out) *status_out = st & ~(int32_t)HVB_STATUS_RERUN_BIT;
    return rc;
}

static int run_stats_host_impl(hvb_plan* pl, const hvb_run_args* a_in, double* stats_out, int32_t* status_out) {
    if (!pl || !a_in) return fail(HVB_ERR_ARG, "null argument");
    const hvb_run_args a_eff = effective_args(pl, a_in);
    const hvb_run_args* a = &a_eff;
    int rc = check_args(pl, a);
    if (rc) return rc;
    if (!stats_out) return fail(HVB_ERR_ARG, "stats_out is NULL");
    const hvb_plan_desc& d = pl->d;
    if (d.ncols <= d.n || (!pl->kernel && pl->kclass != 3)) return fail(HVB_ERR_ARG, "plan has no right-hand sides (dump plan)");
    CU(cudaSetDevice(pl->device));
    const int64_t nF = a->nF, nS = a->nS;
    const int64_t PP = (int64_t)d.P * d.P;
    if (nF == 0 || PP == 0) { if (status_out) *status_out = 0; return HVB_OK; }
    if ((d.n_consts && !a->consts) || (d.n_prefs && !a->prefs)) return fail(HVB_ERR_ARG, "consts/prefs missing");
    if (pl->kclass == 3 && (rc = ensure_gen_kernel(pl, a->prefs, a->consts)) != HVB_OK) return rc;
    cudaStream_t st = pl->streams[0];
    CU(pl->f.ensure((size_t)nF * 8));
    CU(cudaMemcpyAsync(pl->f.p, a->f, (size_t)nF * 8, cudaMemcpyHostToDevice, st));
    if (d.n_samples_cols) {
        CU(pl->samples.ensure((size_t)nS * d.n_samples_cols * 8));
        CU(cudaMemcpyAsync(pl->samples.p, a->samples, (size_t)nS * d.n_samples_cols * 8, cudaMemcpyHostToDevice, st));
    }
    if (a->n_tables) {
        CU(pl->tables.ensure((size_t)a->n_tables * nF * 16));
        CU(cudaMemcpyAsync(pl->tables.p, a->tables, (size_t)a->n_tables * nF * 16, cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemsetAsync(pl->status.p, 0, 4, st));
    rc = push_params(pl, a, st);
    if (rc) return rc;
    // frequency slabs: all samples of a slab stay resident in HBM until they are reduced
    const int64_t budget = (pl->env_stats_mb > 0 ? pl->env_stats_mb : 4096) << 20;
    const int64_t Fc = std::max<int64_t>(std::min<int64_t>(nF, budget / (nS * PP * 16)), 1);
    CU(pl->S[0].ensure((size_t)nS * PP * Fc * 16));
    CU(pl->stats.ensure((size_t)6 * PP * Fc * 8));
    const void* rk = hvb_stats_kernel_ptr();
    for (int64_t fb = 0; fb < nF; fb += Fc) {
        const int64_t nf = std::min(Fc, nF - fb);
        HvbDev D = make_dev(pl);
        D.nF = nf; D.nS = nS; D.ldS = nf; D.pv.table_ld = nF;
        D.f = (const double*)pl->f.p + fb;
        D.samples = d.n_samples_cols ? (const double*)pl->samples.p : nullptr;
        D.pv.tables = a->n_tables ? (const cplx*)pl->tables.p + fb : nullptr;
        D.S = (cplx*)pl->S[0].p;
        D.status = (int*)pl->status.p;
        rc = launch(pl, D, st);
        if (rc) return rc;
        const cplx* dS = (const cplx*)pl->S[0].p;
        long long a_nS = nS, a_rows = PP, a_nF = nf, a_ld = nf, a_ldo = nf;
        double* dout = (double*)pl->stats.p;
        void* args[] = {&dS, &a_nS, &a_rows, &a_nF, &a_ld, &dout, &a_ldo};
        const long long blocks = std::min<long long>(PP * ((nf + 31) / 32), (long long)pl->sm_count * 8);
        CU(cudaLaunchKernel(rk, dim3((unsigned)blocks), dim3(hvb_stats_threads()), args, 0, st));
        pl->launches++;
        CU(cudaMemcpy2DAsync((char*)stats_out + (size_t)fb * 8, (size_t)nF * 8, pl->stats.p, (size_t)nf * 8,
                             (size_t)nf * 8, (size_t)(6 * PP), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    int32_t stw = 0;
    CU(cudaMemcpy(&stw, pl->status.p, 4, cudaMemcpyDeviceToHost));
    if (status_out) *status_out = stw;
    return HVB_OK;
}

extern "C" int hvb_assemble_host(hvb_plan* pl, const hvb_run_args* a_in, double* W_out) {
    if (!pl || !a_in) return fail(HVB_ERR_ARG, "null argument");
    const hvb_run_args a_eff = effective_args(pl, a_in);
    const hvb_run_args* a = &a_eff;
    int rc = check_args(pl, a);
    if (rc) return rc;
    if (!W_out) return fail(HVB_ERR_ARG, "W_out is NULL");
    const hvb_plan_desc& d = pl->d;
    CU(cudaSetDevice(pl->device));
    const int64_t nF = a->nF, nS = a->nS, pts = nF * nS;
    if (pts == 0) return HVB_OK;
    cudaStream_t st = pl->streams[0];
    CU(pl->f.ensure((size_t)nF * 8));
    CU(cudaMemcpyAsync(pl->f.p, a->f, (size_t)nF * 8, cudaMemcpyHostToDevice, st));
    if (d.n_samples_cols) {
        CU(pl->samples.ensure((size_t)nS * d.n_samples_cols * 8));
        CU(cudaMemcpyAsync(pl->samples.p, a->samples, (size_t)nS * d.n_samples_cols * 8, cudaMemcpyHostToDevice, st));
    }
    if (a->n_tables) {
        CU(pl->tables.ensure((size_t)a->n_tables * nF * 16));
        CU(cudaMemcpyAsync(pl->tables.p, a->tables, (size_t)a->n_tables * nF * 16, cudaMemcpyHostToDevice, st));
    }
    rc = push_params(pl, a, st);
    if (rc) return rc;
    const size_t bytes = (size_t)pts * d.n * d.ncols * 16;
    CU(pl->dump.ensure(bytes));
    HvbDev D = make_dev(pl);
    D.nF = nF; D.nS = nS; D.ldS = nF; D.pv.table_ld = nF;
    D.f = (const double*)pl->f.p;
    D.samples = (const double*)pl->samples.p;
    D.pv.tables = (const cplx*)pl->tables.p;
    D.W = (cplx*)pl->dump.p;
    D.region_a = std::max(d.n_dyn, 1);
    D.region_b = std::max(d.coop_scratch, 1);
    const size_t smem = (size_t)(D.region_a + D.region_b) * 16;
    const void* k = hvb_dump_kernel_ptr();
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&D};
    const int grid = (int)std::min<int64_t>(pts, 148 * 16);
    CU(cudaLaunchKernel(k, dim3(grid), dim3(32), args, smem, st));
    pl->launches++;
    CU(cudaMemcpyAsync(W_out, pl->dump.p, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return HVB_OK;
}

// ---- stamp table -> plan, below the ABI (hvb_stamps.cpp) ---------------------------------------------------------------
struct hvb_compiled { StampPlan sp; };

extern "C" int hvb_stamps_compile_host(const hvb_stamp_table* t, int32_t mode, int32_t pad_rows, int32_t pad_rhs, int32_t order,
                                       hvb_compiled** out, hvb_plan_desc* desc_out, const double** consts_out, const int32_t** prefs_out,
                                       int32_t info_out[8], const int32_t** table_prefs_out, const int32_t** table_slots_out) {
    if (!t || !out) return fail(HVB_ERR_ARG, "null argument");
    std::unique_ptr<hvb_compiled> c(new hvb_compiled());
    std::string err;
    if (hvb_stamps_compile(t, mode, pad_rows, pad_rhs, order, &c->sp, &err) != 0) return fail(HVB_ERR_ARG, "stamp table: %s", err.c_str());
    if (desc_out) c->sp.fill_desc(*t, desc_out, 0);
    if (consts_out) *consts_out = c->sp.consts.data();
    if (prefs_out) *prefs_out = c->sp.prefs.data();
    if (info_out) {
        const int32_t info[8] = {c->sp.mode, c->sp.n_real, c->sp.band, c->sp.max_nport, c->sp.n_tables, c->sp.n_sample_cols,
                                 (int32_t)c->sp.table_pref.size(), 0};
        memcpy(info_out, info, sizeof(info));
    }
    if (table_prefs_out) *table_prefs_out = c->sp.table_pref.data();
    if (table_slots_out) *table_slots_out = c->sp.table_slot.data();
    *out = c.release();
    return HVB_OK;
}

extern "C" void hvb_compiled_free(hvb_compiled* c) { delete c; }

extern "C" int hvb_plan_create_from_stamps(const hvb_stamp_table* t, int32_t mode, int device, hvb_plan** out) {
    if (!t || !out) return fail(HVB_ERR_ARG, "null argument");
    std::string err;
    auto create = [&](const StampPlan& sp, int hint, hvb_plan** pl) {
        hvb_plan_desc d;
        sp.fill_desc(*t, &d, hint);
        const int rc = hvb_plan_create(&d, device, pl);
        if (rc == HVB_OK) { (*pl)->def_consts = sp.consts; (*pl)->def_prefs = sp.prefs; }
        return rc;
    };
    const char* nogen = getenv("HVB_NO_GEN");
    const bool gen_on = !(nogen && nogen[0] == '1');
    StampPlan sp;
    // 1. a chain (tridiagonal after the bandwidth-reducing order) -> generated kernel D
    if (gen_on && hvb_stamps_compile(t, mode, 0, 0, HVB_ORDER_RCM, &sp, &err) == 0 && sp.band == 1 && sp.epi_simple && sp.coops.empty()) {
        hvb_plan* pl = nullptr;
        if (create(sp, 2, &pl) == HVB_OK) {
            if (pl->kclass == 3) { *out = pl; return HVB_OK; }
            hvb_plan_destroy(pl);
        }
    }
    // 2. the netlist's own row order: a hub -> generated kernel E; small -> register kernels in an instantiated shape
    if (hvb_stamps_compile(t, mode, 0, 0, HVB_ORDER_NATURAL, &sp, &err) != 0) return fail(HVB_ERR_ARG, "stamp table: %s", err.c_str());
    if (gen_on && sp.coops.size() == HVB_COOP_WORDS && sp.epi_simple && sp.P <= 8) {
        hvb_plan* pl = nullptr;
        if (create(sp, 2, &pl) == HVB_OK) {
            if (pl->kclass == 3 && pl->gen_hub) { *out = pl; return HVB_OK; }
            hvb_plan_destroy(pl);
        }
    }
    int32_t n_pad = 0, p_pad = 0;
    hvb_warp_shape(sp.n_real, sp.P, sp.max_nport, &n_pad, &p_pad);
    if (n_pad > 0) {
        StampPlan padded;
        if (hvb_stamps_compile(t, sp.mode, n_pad, p_pad, HVB_ORDER_NATURAL, &padded, &err) != 0) return fail(HVB_ERR_ARG, "stamp table: %s", err.c_str());
        return create(padded, 1, out);
    }
    // 3. large systems: banded kernel when the ordered matrix is narrow, dense block kernel otherwise
    StampPlan big;
    if (hvb_stamps_compile(t, sp.mode, 0, 0, HVB_ORDER_AUTO, &big, &err) != 0) return fail(HVB_ERR_ARG, "stamp table: %s", err.c_str());
    return create(big, 0, out);
}

extern "C" int hvb_plan_kernel_info(hvb_plan* pl, hvb_kernel_info* out) {
    if (!pl || !out) return fail(HVB_ERR_ARG, "null argument");
    out->kernel_class = pl->kclass;
    out->group = pl->kclass == 0 ? pl->G : pl->NB;       // kernel C: the instantiated half bandwidth
    out->pmax = pl->PMAX;
    out->threads = pl->threads;
    out->smem_bytes = (int32_t)pl->smem;
    out->regs = pl->regs;
    out->ctas_per_sm = pl->ctas_per_sm;
    out->sm_count = pl->sm_count;
    out->rows_per_lane = pl->rows_per_lane;
    out->stream_words = pl->kclass == 2 ? (int32_t)(pl->band_fused ? pl->band_rec_words : (long long)pl->work_words) : 0;
    out->launches = pl->launches;
    out->gen_rolled = pl->gk ? (pl->gk->rolled ? 1 : 0) : -1;
    out->gen_cases = pl->gk ? pl->gk->n_cases : 0;
    out->local_bytes = pl->gk ? pl->gk->k.local_bytes : 0;
    out->gen_scratch_acc = pl->gk ? (pl->gk->scratch_acc ? 1 : 0) : 0;
    out->gen_flops_per_solve = pl->gk ? pl->gk->flops : 0;
    out->gen_hub = pl->gen_hub ? 1 : 0;
    out->gen_two_sided = (pl->gk && pl->gk->two_sided) ? 1 : 0;
    return HVB_OK;
}

// ---- kernel D, host-only introspection (no CUDA call): generated source and an NVRTC compile check ----
extern "C" int hvb_codegen_source(const hvb_plan_desc* desc, const int32_t* prefs, const double* consts, int32_t mode, char* src_out, int64_t src_cap,
                                  int32_t* rows_out, int32_t* itab_out, double* dtab_out, hvb_codegen_info* info) {
    if (!desc || !info) return fail(HVB_ERR_ARG, "null argument");
    memset(info, 0, sizeof(*info));
    GenPlan g;
    g.from_desc(*desc);
    if (consts) g.run_consts.assign(consts, consts + 2 * (size_t)desc->n_consts);
    std::string why, why_hub;
    if (desc->n_prefs && !prefs) return fail(HVB_ERR_ARG, "prefs missing");
    const bool chain = hvb_gen_supported(g, &why);
    const bool hub = !chain && hvb_gen_hub_supported(g, prefs, &why_hub);
    if (!chain && !hub) {
        snprintf(info->reason, sizeof(info->reason), "chain kernel: %s; hub kernel: %s", why.c_str(), why_hub.c_str());
        return HVB_OK;
    }
    GenResult res;
    const int grc = hub ? hvb_gen_hub_source(g, prefs, hvb_gen_prelude(), &res, &why) : hvb_gen_source(g, prefs, mode, hvb_gen_prelude(), &res, &why);
    if (grc != 0) return fail(HVB_ERR_UNSUPPORTED, "kernel generator: %s", why.c_str());
    info->hub = hub ? 1 : 0;
    info->smem_bytes = (int64_t)res.smem_bytes;
    info->threads = res.threads;
    info->supported = 1;
    info->rolled = res.rolled; info->n_cases = res.n_cases; info->scratch_acc = res.scratch_acc;
    info->src_len = (int64_t)res.src.size();
    info->n_rows = (int64_t)res.rows.size(); info->n_itab = (int64_t)res.itab.size(); info->n_dtab = (int64_t)res.dtab.size();
    info->flops_per_solve = res.flops_per_solve;
    info->two_sided = res.two_sided ? 1 : 0; info->cut_row = res.cut_row;
    if (src_out) {
        if (src_cap < (int64_t)res.src.size() + 1) return fail(HVB_ERR_ARG, "source buffer too small");
        memcpy(src_out, res.src.c_str(), res.src.size() + 1);
    }
    if (rows_out) memcpy(rows_out, res.rows.data(), res.rows.size() * 4);
    if (itab_out) memcpy(itab_out, res.itab.data(), res.itab.size() * 4);
    if (dtab_out) memcpy(dtab_out, res.dtab.data(), res.dtab.size() * 8);
    return HVB_OK;
}

extern "C" int hvb_codegen_compile(const char* src, char* log_out, int64_t log_cap, int64_t* cubin_bytes) {
    if (!src) return fail(HVB_ERR_ARG, "null argument");
    std::vector<char> cubin;
    std::string log, err;
    const int rc = hvb_jit_compile(src, &cubin, &log, &err);
    if (log_out && log_cap > 0) snprintf(log_out, (size_t)log_cap, "%s", rc ? err.c_str() : log.c_str());
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    if (rc) return fail(HVB_ERR_CUDA, "NVRTC: %.900s", err.c_str());
    return HVB_OK;
}

extern "C" int hvb_measure_dfma_peak(int device, double* tflops_out) {
    if (!tflops_out) return fail(HVB_ERR_ARG, "null argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* d_out = nullptr;
    CU(cudaMalloc(&d_out, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const void* k = hvb_dfma_probe_ptr();
    int it = iters;
    void* args[] = {&d_out, &it};
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(e0, 0));
        CU(cudaLaunchKernel(k, dim3(blocks), dim3(threads), args, 0, 0));
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;     // 8 chains x 8 unroll FMAs
        if (rep >= 1) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops_out = best;
    return HVB_OK;
}
