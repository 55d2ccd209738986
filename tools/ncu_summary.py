"""Summarise an .ncu-rep: key raw metrics + top stalled SASS instructions.  usage: ncu_summary.py rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
for k, row in enumerate(rows[2:]):
    d = dict(zip(hdr, row))
    print("== launch %d: %s" % (k, d.get("Kernel Name", "?")[:100]))
    for key in KEYS:
        if key in d:
            print("  %-85s %s %s" % (key, d[key], units[hdr.index(key)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        h = i; break
if h is not None:
    hdr = rows[h]; ix = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("== top stalled instructions (of %d samples)" % tot)
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:ntop]
    for i in order:
        r = data[i]
        print("  #%4d %5s  %-60s long_sb=%s short_sb=%s wait=%s lg=%s math=%s exec=%s" % (
            i, r[ix["# Samples"]], r[ix["Source"]][:60], r[ix["stall_long_sb"]], r[ix["stall_short_sb"]],
            r[ix["stall_wait"]], r[ix["stall_lg"]], r[ix["stall_math"]], r[ix["Instructions Executed"]]))
