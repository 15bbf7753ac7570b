"""profiles/traffic.json: measured DRAM traffic per launch of the hot kernels, from `ncu --set full`
captures (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches of a kernel).

    python scripts/ncu_traffic.py <tag> gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]

bench.py reports it as roofline.traffic next to the algorithmic bytes.  ncu flushes the caches before
every replay, so these are cold-cache figures: an upper bound of what a launch moves inside a step,
where its inputs are usually still in the 126 MB L2.
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel function name fragment -> kernel id used by bench.py / advk_kernel_name()
IDS = [("ss_step_bwd_lean_kernel", "ss_step_bwd"), ("ss_step_lean_kernel", "ss_step"), ("ss_step_bwd_box_kernel", "ss_step_bwd"), ("ss_step_bwd_kernel", "ss_step_bwd"), ("ss_step_kernel", "ss_step"), ("smooth3d_xy_kernel<0>", "smooth_fwd_xy"),
       ("smooth3d_z_kernel<0>", "smooth_fwd_z"), ("smooth3d_xy_kernel<1>", "smooth_bwd_xy"),
       ("smooth3d_z_kernel<1>", "smooth_bwd_z"), ("chain_fwd", "chain_fwd"), ("chain_bwd", "chain_bwd"),
       ("init_phi0", "init_phi0"), ("loss_contour_adj", "loss_contour_adj"), ("loss_contour", "loss_contour"),
       ("loss_softmax", "loss_softmax"), ("loss_grad", "loss_grad"), ("adjoint_axis_big", "adjoint_axis")]


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    acc = collections.OrderedDict()
    for rep in reps:
        raw = (open(rep).read() if rep.endswith(".csv") else
                   subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        kn, ir, iw, it = (hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"),
                          hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"))
        for r in rows[2:]:
            kid = next((k for frag, k in IDS if frag in r[kn]), None)
            if kid is None:
                continue
            b = float(r[ir].replace(",", "")) * unit_scale(units[ir]) + float(r[iw].replace(",", "")) * unit_scale(units[iw])
            t = float(r[it].replace(",", "")) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[it], 1.0)
            e = acc.setdefault(kid, {"n": 0, "bytes": 0.0, "us": 0.0, "kernel": r[kn][:80]})
            e["n"] += 1
            e["bytes"] += b
            e["us"] += t
    out = collections.OrderedDict()
    for k, e in acc.items():
        out[k] = {"dram_bytes_per_launch": e["bytes"] / e["n"], "launches_captured": e["n"],
                  "ncu_us_per_launch": e["us"] / e["n"], "kernel": e["kernel"], "capture": tag,
                  "workload": "m128 (1x1x128^3)"}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
