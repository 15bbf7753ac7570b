"""Compact per-launch table of the interesting ncu metrics from a raw CSV page.
    ncu -i rep --page raw --csv > raw.csv ; python scripts/ncu_table.py raw.csv"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = [("us", "gpu__time_duration.sum"), ("dR_MB", "dram__bytes_read.sum"), ("dW_MB", "dram__bytes_write.sum"),
        ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("Minst", "smsp__inst_executed.sum"), ("l1hit", "l1tex__t_sector_hit_rate.pct"), ("l2hit", "lts__t_sector_hit_rate.pct"),
        ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("red_req", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum"), ("grid", "launch__grid_size")]
idx = [(n, hdr.index(w)) for n, w in want if w in hdr]
kn = hdr.index("Kernel Name")
print("kernel | " + " | ".join(n for n, _ in idx))
for r in rows[2:]:
    vals = []
    for n, i in idx:
        try:
            v = float(r[i].replace(",", ""))
            if n == "Minst":
                v /= 1e6
            vals.append("%.4g" % v)
        except ValueError:
            vals.append(r[i])
    print(r[kn][:46] + " | " + " | ".join(vals))
