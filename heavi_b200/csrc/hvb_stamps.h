// hvb_stamps.h -- stamp-table compiler (hvb_stamps.cpp): host-only, no CUDA.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/heavi_b200.h"

struct StampPlan {                     // what plan.py calls DevicePlan: the arrays of hvb_plan_desc + per-run defaults
    int mode = 0, n = 0, ncols = 0, P = 0, n_real = 0, band = -1, n_dyn = 0, coop_scratch = 0, max_nport = 0;
    int n_tables = 0, n_sample_cols = 0, epi_simple = 0;
    std::vector<double> consts, const_vals, fop_arg, epi_scale;
    std::vector<int32_t> prefs, ops, coops, cell_ptr, cell_ent, row_ptr, row_ent, fop_code, fop_out, epi_port_of_row, port_rows, port_zpref, table_pref, table_slot;
    void fill_desc(const hvb_stamp_table& t, hvb_plan_desc* d, int kernel_hint) const;
};

bool hvb_stamps_norton_legal(const hvb_stamp_table* t);
// pad_rows == 0: no padding.  order: HVB_ORDER_*.
int hvb_stamps_compile(const hvb_stamp_table* t, int mode, int pad_rows, int pad_rhs, int order, StampPlan* out, std::string* err);
