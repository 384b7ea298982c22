"""Summarise an .ncu-rep: key throughput metrics and warp-stall breakdown per captured launch.
    python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            print("%-72s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
    if not st:
        st = [(float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith(".pct") and r[i]]
    for v, h in sorted(st, reverse=True)[:8]:
        print("   stall %-60s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__warp_issue_stalled_", ""), v))
