"""profiles/traffic.json: measured DRAM traffic per launch of our kernels, from `ncu --set full` captures
(dram__bytes_read.sum + dram__bytes_write.sum), keyed by the FULL TEMPLATE INSTANCE of the kernel.

    python scripts/ncu_traffic.py <tag> gpurun_out/a_raw.csv|a.ncu-rep [...]

`instances`  : one entry per kernel instance (e.g. "lean_warp_bwd_pk_kernel<3, 0, 1>"), mean over its captured
               launches.
`by_bench_id`: the same grouped by the kernel id bench.py / advk_kernel_name() uses; when an id covers several
               instances (ss_step_bwd: the first adjoint level does not zero its upstream buffer) the figure is
               the launch-weighted mean of ONE captured iteration and the instances are listed -- never a mean
               of unrelated kernels.
bench.py reports by_bench_id[...] as roofline.traffic next to the algorithmic bytes.  ncu flushes the caches
before every replay, so these are cold-cache figures: an upper bound of what a launch moves inside a step.
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel function name -> kernel id of ADVK_KERNELS (advk_common.cuh); first match wins
IDS = [("ss_step_bwd_", "ss_step_bwd"), ("ss_step_", "ss_step"), ("smooth2d_kernel<0>", "smooth_fwd"),
       ("smooth2d_kernel<1>", "smooth_bwd"), ("smooth3d_xy_kernel<0>", "smooth_fwd"), ("smooth3d_z_kernel<0>", "smooth_fwd"),
       ("smooth3d_xy_kernel<1>", "smooth_bwd"), ("smooth3d_z_kernel<1>", "smooth_bwd"),
       ("smooth3d_tma_kernel<0>", "smooth_fwd"), ("smooth3d_tma_kernel<1>", "smooth_bwd"),
       ("init_phi0", "init_phi0"), ("loss_contour_adj", "loss_contour_adj"), ("loss_contour", "loss_contour"),
       ("loss_softmax", "loss_softmax"), ("loss_grad", "loss_grad"), ("adjoint_axis_f", "adjoint_axis_f"),
       ("adjoint_axis", "adjoint_axis"), ("lowres_smooth", "lowres_smooth"), ("aos_to_planar", "aos_to_planar"),
       ("lean_intensity_kernel<2, false>", "chain_img_fwd"), ("lean_intensity_kernel<3, false>", "chain_img_fwd"),
       ("lean_intensity_kernel<2, true>", "chain_img_bwd"), ("lean_intensity_kernel<3, true>", "chain_img_bwd"),
       ("lean_intensity_kernel<2, 0>", "chain_img_fwd"), ("lean_intensity_kernel<3, 0>", "chain_img_fwd"),
       ("lean_intensity_kernel<2, 1>", "chain_img_bwd"), ("lean_intensity_kernel<3, 1>", "chain_img_bwd"),
       ("lean_warp_fwd_pk", "chain_pk_fwd"), ("lean_pack", "chain_pk_fwd"), ("lean_warp_bwd_pk", "chain_pk_bwd"),
       ("lean_unpack", "chain_pk_bwd"), ("lean_warp_fwd", "chain_img_fwd"), ("lean_warp_bwd", "chain_img_bwd"),
       ("lean_affine", "chain_img_fwd"),
       ("chain_fwd", "chain_fwd"), ("chain_bwd", "chain_bwd"), ("update_kernel", "update"), ("sumsq", "sumsq"),
       ("lowfield_fwd", "lowfield_fwd"), ("lowfield_bwd", "lowfield_bwd"), ("affine_theta_fwd", "affine_theta_fwd"),
       ("affine_theta_bwd", "affine_theta_bwd"), ("loss_finalize", "loss_finalize")]


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def instance_name(full):
    s = re.sub(r"^void\s+", "", full)
    s = re.sub(r"\(.*$", "", s)
    return re.sub(r"^advk::", "", s)


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    inst = collections.OrderedDict()
    for rep in reps:
        raw = (open(rep).read() if rep.endswith(".csv") else
               subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        kn, ir, iw, it = (hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"),
                          hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"))
        for r in rows[2:]:
            name = instance_name(r[kn])
            kid = next((k for frag, k in IDS if frag in name), None)
            if kid is None:
                continue
            b = float(r[ir].replace(",", "")) * unit_scale(units[ir]) + float(r[iw].replace(",", "")) * unit_scale(units[iw])
            t = float(r[it].replace(",", "")) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[it], 1.0)
            e = inst.setdefault(name, {"n": 0, "bytes": 0.0, "us": 0.0, "bench_id": kid})
            e["n"] += 1
            e["bytes"] += b
            e["us"] += t
    instances = collections.OrderedDict()
    by_id = collections.OrderedDict()
    for name, e in inst.items():
        instances[name] = {"dram_bytes_per_launch": e["bytes"] / e["n"], "launches_captured": e["n"],
                           "ncu_us_per_launch": e["us"] / e["n"], "bench_id": e["bench_id"]}
        g = by_id.setdefault(e["bench_id"], {"n": 0, "bytes": 0.0, "us": 0.0, "instances": []})
        g["n"] += e["n"]
        g["bytes"] += e["bytes"]
        g["us"] += e["us"]
        g["instances"].append(name)
    out = {"capture": tag, "workload": "m128 (1x1x128^3), one PGD iteration", "instances": instances,
           "by_bench_id": collections.OrderedDict(
               (k, {"dram_bytes_per_launch": g["bytes"] / g["n"], "launches_captured": g["n"],
                    "ncu_us_per_launch": g["us"] / g["n"], "ncu_us_per_step": g["us"], "instances": g["instances"]})
               for k, g in by_id.items())}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["by_bench_id"], indent=1))


if __name__ == "__main__":
    main()
