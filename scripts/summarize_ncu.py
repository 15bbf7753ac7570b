"""Turn the raw ncu outputs of a GPU visit (gpurun_out/<tag>_launches.csv, <tag>_prof.ncu-rep) into
the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag> [<out-prefix>]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def launches():
    path = os.path.join(G, tag + "_launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot, ours = collections.OrderedDict(), 0.0, 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        if "advk::" in name:
            ours += v
    with open(os.path.join(P, prefix + "_launch_list.md"), "w") as f:
        f.write("# ncu launch list -- one PGD inner-loop iteration (scripts/one_step.py, workload m128)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off`; "
                "per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("launches: %d, total %.1f us, advk kernels %.1f us (%.1f %%)\n\n" % (
            len(data), tot, ours, 100 * ours / max(tot, 1e-9)))
        f.write("| launches | total us | share | kernel |\n|---:|---:|---:|---|\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %d | %.1f | %.1f %% | `%s` |\n" % (c, t, 100 * t / tot, name[:110]))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def full():
    import glob
    reps = sorted(glob.glob(os.path.join(G, tag + "_full_*.ncu-rep")) + glob.glob(os.path.join(G, tag + "_full_*.csv")))
    legacy = os.path.join(G, tag + "_prof.ncu-rep")
    if os.path.exists(legacy):
        reps.append(legacy)
    if not reps:
        return
    with open(os.path.join(P, prefix + "_ncu_full.md"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (per launch; cold cache, serialised) -- %s\n\n" % tag)
        f.write("Workload m128 (1x1x128^3, full chain), `scripts/one_step.py`; one block per captured launch.\n\n")
        for rep in reps:
            raw = (open(rep).read() if rep.endswith(".csv") else
                   subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
            rows = list(csv.reader(raw.splitlines()))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            idx = [(w, hdr.index(w)) for w in WANT + EXTRA if w in hdr]
            kn = hdr.index("Kernel Name")
            for r in rows[2:]:
                f.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % r[kn][:100])
                for w, i in idx:
                    f.write("| %s | %s | %s |\n" % (w, r[i], units[i]))
                f.write("\n")
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_traffic.py"), tag] + reps)


EXTRA = ["l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
         "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
         "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "smsp__warps_eligible.avg.per_cycle_active"]

launches()
full()
print("wrote profiles/%s_*" % prefix)
