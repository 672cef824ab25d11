"""Selected metrics of an .ncu-rep as CSV: python tools/ncu_summary.py report.ncu-rep > profiles/x.csv"""
import csv, subprocess, sys
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "stalled_barrier_per_issue", "stalled_long_scoreboard_per_issue", "stalled_short_scoreboard_per_issue",
        "stalled_mio_throttle_per_issue", "stalled_lg_throttle_per_issue", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "lts__t_sector_hit_rate.pct")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
keep = [i for i, c in enumerate(h) if c in ("ID", "Kernel Name") or any(k in c for k in KEEP)]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] for i in keep])
