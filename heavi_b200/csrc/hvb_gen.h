// hvb_gen.h -- kernel D: netlist-specialised CUDA source for chain-like (tridiagonal) systems.
//
// Host-only (no CUDA call in here): turns a frozen plan into the source text of ONE kernel whose
// control flow, value formulas, matrix cells and port positions are those of this netlist, so the
// device executes no plan interpreter at all.  hvb_jit.cpp compiles the text with NVRTC for
// sm_100a; hvb_api.cu launches it (kernel class 3).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/heavi_b200.h"

struct GenPlan {                       // host copy of what the generator reads from hvb_plan_desc
    int n = 0, ncols = 0, P = 0, nvc = 0, ndyn = 0, nops = 0, nfast = 0, ncoops = 0, epi_simple = 0, nprefs = 0;
    std::vector<double> const_vals;    // [2*nvc]
    std::vector<int32_t> ops;          // [nops*8]
    std::vector<int32_t> row_ptr, row_ent;
    std::vector<int32_t> fop_code, fop_out;
    std::vector<double> fop_arg;
    std::vector<int32_t> port_rows, epi_port_of_row;
    std::vector<double> epi_scale;
    std::vector<double> run_consts;    // optional [2*n_consts]: the run's constants; when present, constant parameters are
                                       // baked into the code (literals / constant bank) instead of being loaded from HBM
    std::vector<int32_t> coops;        // [ncoops*8]
    std::vector<int32_t> spl_trace;    // [n_traces*4]
    bool allow_two_sided = true;       // false keeps a long chain on the one-sided row loop (modes 0 and 3)
    int occ_class = 0;                 // two-sided row loop: 0 = the compiler's own register allocation, 1 = 128 registers
                                       // (512 threads per SM), 2 = capped at 168 registers (384 threads per SM)
    int cut_col = -1;                  // generator-internal (half chains of the two-sided form): column index under which the
                                       // last row's coupling to the other half is filed; -1 in a plan that came from a desc
    void from_desc(const hvb_plan_desc& d);
};

struct GenResult {
    std::string src;                   // CUDA C++ translation unit (prelude + kernel)
    std::string kernel_name;
    bool rolled = false;               // true: loop over rows + shape switch, constants in the tables below
    std::vector<int32_t> rows;         // rolled: [n-1][4] assembly shape, step shape, itab base, dtab base
    std::vector<int32_t> itab;
    std::vector<double> dtab;
    bool scratch_acc = false;          // port-pair accumulators live in the output array, per-port sums in shared memory (P > 4)
    bool tables_in_source = false;     // rolled: the row tables are __constant__ initialisers of the module (no upload needed)
    int threads = 128;                 // threads per CTA the kernel is written for
    size_t smem_bytes = 0;             // dynamic shared memory per CTA (kernel E: the per-thread dense tiles)
    int n_cases = 0;                   // distinct row shapes emitted (rolled)
    bool two_sided = false;            // rolled, walked from both ends towards a cut row (sets HVB_STATUS_RERUN_BIT when unsafe)
    int cut_row = -1;                  // two-sided: last row of the front half
    long long flops_per_solve = 0;     // executed real flops of the elimination / border recurrences (excl. stamp values)
};

// why == nullptr or receives the reason when the plan cannot be served by kernel D
bool hvb_gen_supported(const GenPlan& g, std::string* why);

// prefs: the run's parameter references [nprefs][2] (kind, index): value formulas are specialised on the kinds.
// mode: 0 auto (unrolled up to HVB_GEN_UNROLL_MAX rows; above that rolled, two-sided when that saves work),
//       1 unrolled, 2 rolled one-sided, 3 rolled two-sided (falls back to 2 when the chain offers no cut).
// prelude: text of plan_format.h + hvb_device.cuh (embedded at build time, see Makefile).
int hvb_gen_source(const GenPlan& g, const int32_t* prefs, int mode, const char* prelude, GenResult* out, std::string* err);

// Kernel E ("hub": one N-port S block + chains hanging off its nodes, hvb_gen_hub.cpp)
bool hvb_gen_hub_candidate(const GenPlan& g);                                   // structure only
bool hvb_gen_hub_supported(const GenPlan& g, const int32_t* prefs, std::string* why);
int hvb_gen_hub_source(const GenPlan& g, const int32_t* prefs, const char* prelude, GenResult* out, std::string* err);

const char* hvb_gen_prelude();
