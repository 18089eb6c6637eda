"""Print the handful of ncu metrics we track from a .ncu-rep (first kernel in the report)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:60])
    for h, u, v in zip(hdr, units, vals):
        if any(h == k or h.startswith(k) for k in KEYS) and "per_second" not in h and "pct_of_peak_sustained_elapsed" not in h or "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            print("  %-75s %-10s %s" % (h, u, v))
