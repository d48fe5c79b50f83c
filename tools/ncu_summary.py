"""Summarise an .ncu-rep (key metrics + top stall lines). usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
f = sys.argv[1]
raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active',
 'launch__registers_per_thread','launch__occupancy_limit_registers','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg','sm__cycles_elapsed.avg',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print("%-95s %s %s" % (k, r[i], units[i]))
    print("---")
if "--source" in sys.argv:
    src = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    sh = srows[0]
    def col(name):
        for i, h in enumerate(sh):
            if h.strip() == name: return i
        return None
    ci = col("Source"); cs = col("# Samples") or col("Samples"); cw = col("Warp Stall Sampling (All Samples)")
    print(sh[:12])
    c = cw if cw is not None else cs
    if c is not None:
        top = sorted([r for r in srows[1:] if len(r) > c and r[c].replace('.','',1).isdigit()], key=lambda r: -float(r[c]))[:25]
        for r in top: print(r[c], '|', r[ci][:120] if ci is not None else r[:3])
