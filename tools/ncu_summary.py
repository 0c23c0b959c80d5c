"""Summarise an .ncu-rep (CPU side): headline metrics + executed-instruction histogram over the SASS.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [bucket]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1]))
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__cycles_elapsed.max"]
for k in want:
    if k in d: print("%-86s %s" % (k, d[k]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iN, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[iS].strip(), int(r[iN]), int(r[iSm])) for r in rows[2:] if len(r) > iN and r[iN].isdigit()]
tot = sum(x[1] for x in ins) or 1; tots = sum(x[2] for x in ins) or 1
print("SASS instructions %d, executed %d, samples %d" % (len(ins), tot, tots))
byop = collections.Counter()
for s, n, _ in ins:
    byop[(s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]] += n
print(" ".join("%s:%.1f%%" % (k, 100 * v / tot) for k, v in byop.most_common(16)))
for b in range(0, len(ins), B):
    seg = ins[b:b + B]; n = sum(x[1] for x in seg); sm = sum(x[2] for x in seg)
    if n == 0: continue
    ops = collections.Counter()
    for s, c, _ in seg:
        ops[(s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]] += c
    print("%5d-%5d inst %5.1f%% samples %5.1f%% | %s" % (b, b + B, 100 * n / tot, 100 * sm / tots, " ".join("%s:%.0fk" % (k, v / 1e3) for k, v in ops.most_common(7))))
