#!/usr/bin/env python
"""bench.py -- adversarial inner-loop iterations/sec on 3-D 128^3 volumes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one PGD inner-loop iteration (adv_compose_solver.py:308-368 of the reference):
chain forward (noise -> bias -> morph -> affine), model forward, warp-back of the K-channel
prediction, valid-region mask, consistency loss, backward to the four parameter sets and their
PGD updates.  The transforms run in libadvchain_b200.so (sm_100a CUDA through the C ABI); the
toy model `Conv3d(1,4,3,1,1)` and the consistency loss are PyTorch (out of the path's scope,
SURVEY.md section 8) but are inside the timed step because the inner loop cannot run without them.

Multi-GPU: one process per GPU (torchrun), every rank owns its own volumes (batch sharding, no
data-path collective) -> weak scaling; value = volume-iterations of all ranks /
max-over-ranks time.

`--impl reference` times the reference's OWN code (the unmodified package mirrored under baseline/_ref by
`__graft_entry__.build()`; kind "reference") on the box's host cores at the workload's FULL size -- no voxel
scaling; as many of the K steps as fit a 150 s budget, at least one, reported as `steps` -- rank 0 only.  If
baseline/_ref is absent the CPU restatement (oracle/, kind "port") is timed instead.
`--impl reference-cuda` runs the same unmodified reference with device=cuda on the same B200 (PyTorch
eager: the incumbent users run today); the b200 line carries that number as `cuda_eager_baseline`.

Workloads: m128 (the metric; weak scaling, one 128^3 volume per GPU), c2 / c3 (BASELINE configs 2 and 3 on one
GPU), c4 / c5 (BASELINE configs 4 and 5: FIXED global batch 256 / 16 split over the ranks -> strong scaling;
`--exact-global` adds the three scalar all-reduces that make the shards reproduce the unsharded run).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "adv inner-loop iters/sec on 3D 128^3 volumes"
UNIT = "iters/s"

WORKLOADS = {
    # name: (d, per-GPU size, chain)
    "m128": (3, [1, 1, 128, 128, 128], ["noise", "bias", "morph", "affine"]),
    "c3": (3, [2, 1, 128, 128, 64], ["noise", "bias", "morph", "affine"]),
    "c2": (2, [8, 1, 256, 256], ["noise", "bias", "morph", "affine"]),
    "c4shard": (2, [32, 1, 256, 256], ["noise", "bias", "morph", "affine"]),
    "c5shard": (3, [2, 1, 256, 256, 128], ["morph", "affine"]),
    "tiny3d": (3, [1, 1, 32, 32, 32], ["noise", "bias", "morph", "affine"]),
    # strong scaling: the size is the GLOBAL batch, split over the ranks (BASELINE.json configs 4 and 5)
    "c4": (2, [256, 1, 256, 256], ["noise", "bias", "morph", "affine"]),
    "c5": (3, [16, 1, 256, 256, 128], ["morph", "affine"]),
}
STRONG = ("c4", "c5")
K_CLASSES = 4

# Algorithmic 4-byte words per voxel moved by ONE launch of a kernel (SURVEY.md section 8d / DESIGN.md
# "Kernels"): compulsory reads + writes with perfect reuse inside the launch.  d = spatial dims.
ALGO_WORDS = {
    "ss_step": lambda d: 2 * d,          # R phi_{k-1} (gather source == grid), W phi_k
    "ss_step_bwd": lambda d: 3 * d,      # R phi_{k-1}, R g_k, W g_{k-1}
    "ss_fused": lambda d: 2 * d,
    "smooth_fwd": lambda d: 3 * d,       # R phi_n, R phi_0, W field
    "smooth_bwd": lambda d: 5 * d,       # R g_field, R field, R phi_n, R phi_0, W g_off
    "init_phi0": lambda d: d,
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--exact-global", action="store_true",
                    help="multi-GPU: ShardContext (scalar all-reduces) so that the shards reproduce the unsharded run")
    ap.add_argument("--no-cuda-baseline", action="store_true", help="skip the reference-on-CUDA (PyTorch eager) leg")
    ap.add_argument("--workload", default="m128", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the PGD loop eagerly (no CUDA graph)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel breakdown JSON here")
    return ap.parse_args()


# ----------------------------------------------------------------------------------- helpers

def make_cfgs(d, size):
    from tests.golden.cases import stage_cfgs
    return stage_cfgs(d, size)


class ClockSampler(object):
    """nvidia-smi clock / throttle-reason sampler running beside the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_ready(self, timeout_s=15.0):
        """Blocks until the first sample has been written (the first nvidia-smi call on a fresh box can
        take seconds to come up)."""
        if self.p is None:
            return False
        t0 = time.time()
        while time.time() - t0 < timeout_s:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return True
            except OSError:
                return False
            time.sleep(0.05)
        return False

    @staticmethod
    def _epoch(ts):
        """nvidia-smi timestamp 'YYYY/MM/DD HH:MM:SS.mmm' (local time) -> epoch seconds."""
        try:
            base, _, ms = ts.partition(".")
            return time.mktime(time.strptime(base, "%Y/%m/%d %H:%M:%S")) + (float("0." + ms) if ms else 0.0)
        except Exception:
            return None

    def stop(self, window=None):
        """window = (t0, t1) wall-clock bounds of the timed region: only samples taken inside it (the
        samples 'under load') enter the median; without any, every sample is used and `in_window` is 0."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                rows.append((self._epoch(c[0]), float(c[1]), float(c[2]),
                             [nm for nm, v in zip(names, c[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = []
        if window is not None:
            inside = [r for r in rows if r[0] is not None and window[0] - 0.03 <= r[0] <= window[1] + 0.03]
            if not inside:                      # a timed region shorter than the sampling period
                mid = 0.5 * (window[0] + window[1])
                near = sorted((abs(r[0] - mid), i) for i, r in enumerate(rows) if r[0] is not None)
                inside = [rows[i] for dist_, i in near[:1] if dist_ < 0.5]
        use = inside or rows
        sm = [r[1] for r in use]
        mx = [r[2] for r in use]
        reasons = set(nm for r in use for nm in r[3])
        out["in_window"] = len(inside)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # under load = upper half of the samples (the sampler also sees the idle edges)
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def ncu_traffic_bytes(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture (profiles/traffic.json, written by scripts/summarize_ncu.py), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        e = t.get("by_bench_id", t).get(kernel)
        return None if e is None else float(e["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------- CPU arm

def cpu_port_step_time(d, size, chain, steps=1, warmup=0, threads=None):
    """Times `steps` PGD inner-loop iterations of the CPU port (oracle/) on `size`; returns the
    list of per-step seconds."""
    from oracle import advchain_oracle as orc
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfgs = make_cfgs(d, size)
    stages = [orc.make_stage(n, cfgs[n]) for n in chain]
    sol = orc.Solver(stages, if_norm_image=True, min_intensity=0.0, max_intensity=1.0)
    torch.manual_seed(0)
    data = torch.rand(*size)
    conv = torch.nn.Conv2d if d == 2 else torch.nn.Conv3d
    model = conv(size[1], K_CLASSES, 3, 1, 1).eval()
    for s in stages:
        s.init()
    with torch.no_grad():
        init_out = model(data)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        sol.inner_loop(model, data, init_out, 1)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def _nvox(size):
    n = size[0]
    for s in size[2:]:
        n *= s
    return n


REF_BUDGET_S = 150.0     # wall-clock bound of the reference arm's timed region


def load_reference():
    """The UNMODIFIED reference package (baseline/_ref, mirrored by __graft_entry__.install_reference(); in the
    build container /root/reference itself), or None.  Test/bench infrastructure only: the product path never
    imports it."""
    try:
        from oracle import ref_shim
        if ref_shim.available():
            return ref_shim.load()
    except Exception as exc:                                  # pragma: no cover
        sys.stderr.write("bench.py: reference not loadable (%s)\n" % (exc,))
    return None


def reference_stepper(aug, d, size, chain, device):
    """One PGD inner-loop iteration of the reference's own solver (adv_compose_solver.py:289-405) on
    synthetic data of the workload's shape; returns step()."""
    cfgs = make_cfgs(d, size)
    use_gpu = device.type == "cuda"
    ts = []
    for n in chain:
        cls = {"noise": aug.AdvNoise, "bias": aug.AdvBias, "morph": aug.AdvMorph, "affine": aug.AdvAffine}[n]
        ts.append(cls(spatial_dims=d, config_dict=cfgs[n], use_gpu=use_gpu, device=device))
    sol = aug.ComposeAdversarialTransformSolver(
        chain_of_transforms=ts, divergence_types=["mse", "contour"], divergence_weights=[1.0, 0.5],
        use_gpu=use_gpu, if_norm_image=True, min_intensity=0.0, max_intensity=1.0)
    torch.manual_seed(0)
    data = torch.rand(*size).to(device)
    conv = torch.nn.Conv2d if d == 2 else torch.nn.Conv3d
    model = conv(size[1], K_CLASSES, 3, 1, 1).eval().to(device)
    init_out = sol.get_init_output(model=model, data=data)
    sol.init_random_transformation()
    flags, steps = [True] * len(chain), [1.0] * len(chain)

    def step():
        sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags,
                                 n_iter=1, step_sizes=steps)
    return step


def small_size(d, size):
    return [min(size[0], 2), size[1]] + [32] * d


def time_cpu_arm(d, size, chain, steps, budget_s, prefer_reference=True):
    """Times up to `steps` PGD iterations at FULL size on the host cores within `budget_s` (at least one).
    -> (seconds per timed step list, kind, cores)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    aug = load_reference() if prefer_reference else None
    cpu = torch.device("cpu")
    if aug is not None:
        kind = "reference"
        warm = reference_stepper(aug, d, small_size(d, size), chain, cpu)
        step = reference_stepper(aug, d, size, chain, cpu)
    else:
        kind = "port"
        warm = lambda: cpu_port_step_time(d, small_size(d, size), chain, steps=1)     # noqa: E731
        step = None
    warm()                               # thread pools, lazy imports, oneDNN primitives: on a small volume
    times = []
    t_start = time.perf_counter()
    while len(times) < max(1, steps):
        t0 = time.perf_counter()
        if step is not None:
            step()
        else:
            cpu_port_step_time(d, size, chain, steps=1)
        times.append(time.perf_counter() - t0)
        spent = time.perf_counter() - t_start
        if spent + max(times) > budget_s:
            break
    return times, kind, cores


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path at the workload's full size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    d, size, chain = WORKLOADS[args.workload]
    times, kind, cores = time_cpu_arm(d, size, chain, args.steps, REF_BUDGET_S)
    n = len(times)
    total = sum(times)
    value = n / total
    sample = ("%d PGD inner-loop iteration(s) of the %s on %s at FULL size (no voxel scaling), %d host threads; "
              "%d of the %d requested steps fit the %.0f s budget; warm-up = 1 iteration on a %s volume"
              % (n, "unmodified reference (baseline/_ref, CPU)" if kind == "reference" else "CPU port (oracle/)",
                 "x".join(map(str, size)), cores, n, args.steps, REF_BUDGET_S, "x".join(map(str, small_size(d, size)))))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": n, "steps_requested": args.steps, "warmup": 1, "ms_per_step": 1e3 * total / n,
        "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, d, size, chain),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def time_reference_cuda(d, size, chain, dev, steps=5, warmup=3):
    """The unmodified reference with device=cuda (PyTorch eager on the same GPU). -> (ms per iteration, detail) or
    (None, None).  Timed under both settings of torch.backends.cudnn.benchmark (the reference's Gaussian and the toy
    model are cuDNN convolutions: PyTorch's default is False, this file sets True for the user model) and the
    FASTER one is reported as the incumbent; one run measured 146 ms and another 1 356 ms per iteration with the
    flag on (gpurun_out/r02f, r02q), so a single setting is not a fair baseline."""
    aug = load_reference()
    if aug is None:
        return None, None
    saved = torch.backends.cudnn.benchmark
    detail = {}
    try:
        for flag in (False, True):
            torch.backends.cudnn.benchmark = flag
            step = reference_stepper(aug, d, size, chain, dev)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            detail["cudnn_benchmark_%s" % flag] = e0.elapsed_time(e1) / steps
    finally:
        torch.backends.cudnn.benchmark = saved
    return min(detail.values()), detail


def run_reference_cuda(args):
    """`--impl reference-cuda`: the incumbent -- the reference's PyTorch-eager path on the same B200."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    d, size, chain = WORKLOADS[args.workload]
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-cuda", "unavailable": "no CUDA device"}))
        return 0
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    steps = max(1, min(args.steps, 20))
    ms, detail = time_reference_cuda(d, size, chain, dev, steps=steps, warmup=max(1, min(args.warmup, 3)))
    if ms is None:
        print(json.dumps({"impl": "reference-cuda", "unavailable": "baseline/_ref (mirror of the reference package) "
                          "is not present; run __graft_entry__.build() where /root/reference exists"}))
        return 0
    print(json.dumps({
        "impl": "reference-cuda", "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": 1, "steps": steps,
        "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, d, size, chain),
        "kind": "unmodified reference (baseline/_ref), device=cuda, PyTorch eager",
        "ms_per_step_by_setting": detail}))
    return 0


def workload_config(args, d, size, chain, per_gpu=None):
    strong = args.workload in STRONG
    per_gpu = per_gpu or size
    cfg = {"workload": "%s: %dD %s %s, chain %s, 1 PGD step per iteration, toy model Conv%dd(1,%d,3,1,1), "
                       "loss mse+contour" % (args.workload, d, "x".join(map(str, size)),
                                             "GLOBAL batch split over the GPUs" if strong else "per GPU",
                                             "->".join(chain), d, K_CLASSES),
           "per_gpu_size": per_gpu, "chain": chain, "n_gpus": args.gpus,
           "l2": "no flush between iterations: the per-step working set (field levels alone: 2 signs x 10 fields x %d "
                 "B/voxel = %.0f MB per GPU) %s the 126 MB L2"
                 % (16 if d == 3 else 8, 2 * 10 * (16 if d == 3 else 8) * _nvox(per_gpu) / 1e6,
                    "exceeds" if 2 * 10 * (16 if d == 3 else 8) * _nvox(per_gpu) / 1e6 > 126 else "FITS IN (not a valid timing size)"),
           "model_precision": "the toy model runs under PyTorch defaults (cuDNN may pick TF32 convolution "
                              "kernels); every advk kernel computes in fp32"}
    if strong:
        cfg["global_batch"] = size[0]
        cfg["exact_global"] = bool(getattr(args, "exact_global", False))
    return cfg


# ----------------------------------------------------------------------------------- GPU arm

def _release_graphs():
    """Captured iterations of an exact-global run hold NCCL work: they go before the process group does."""
    from advchain_b200.augmentor.solver import release_graphs
    release_graphs()


def build_solver(d, size, chain, dev):
    from advchain_b200.augmentor import (AdvAffine, AdvBias, AdvMorph, AdvNoise,
                                         ComposeAdversarialTransformSolver)
    cfgs = make_cfgs(d, size)
    ts = []
    for n in chain:
        cls = {"noise": AdvNoise, "bias": AdvBias, "morph": AdvMorph, "affine": AdvAffine}[n]
        ts.append(cls(d, cfgs[n], device=dev))
    return ComposeAdversarialTransformSolver(
        ts, divergence_types=["mse", "contour"], divergence_weights=[1.0, 0.5], if_norm_image=True,
        min_intensity=0.0, max_intensity=1.0)


def run_b200(args):
    import torch.distributed as dist
    from advchain_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    _lib.load()   # fail loudly if the CUDA library is missing
    torch.backends.cudnn.benchmark = True      # the user model's convolutions: let cuDNN pick its plan
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d, gsize, chain = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    size = list(gsize)
    ctx = None
    if strong:
        from advchain_b200.augmentor.sharding import ShardContext, shard_slice
        sl = shard_slice(gsize[0], rank, world)
        size[0] = sl.stop - sl.start
        if args.exact_global and world > 1:
            ctx = ShardContext(gsize[0])
    sampler = ClockSampler(local) if rank == 0 else None     # runs beside everything; filtered to the timed window
    torch.manual_seed(1234 + rank)
    host_data = torch.rand(*size).pin_memory()
    conv = torch.nn.Conv2d if d == 2 else torch.nn.Conv3d
    torch.manual_seed(0)
    model = conv(size[1], K_CLASSES, 3, 1, 1).eval().to(dev)
    for p in model.parameters():
        p.requires_grad_(True)
    sol = build_solver(d, size, chain, dev)
    sol.shard = ctx
    data = host_data.to(dev)
    init_out = sol.get_init_output(model, data)
    sol.init_random_transformation()
    flags = [True] * len(chain)
    steps_sz = [1.0] * len(chain)
    host_loss = torch.zeros(1).pin_memory()

    # end-to-end leg: the volume of every step comes from pinned host memory and the step's loss goes
    # back to the host.  The copies are pipelined the way a training loop would do it: the upload of step
    # i+1 runs on a copy stream into the other half of a double buffer while step i computes, and the
    # host reads the loss of step i-1 (its own event) while step i is in flight -- every step's bytes
    # still cross PCIe inside the timed region, but the compute stream never waits for the host.
    copy_stream = torch.cuda.Stream()
    dbuf = [torch.empty_like(data), torch.empty_like(data)]
    up_evt = [torch.cuda.Event(), torch.cuda.Event()]
    free_evt = [torch.cuda.Event(), torch.cuda.Event()]
    host_loss2 = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
    loss_dev = [torch.zeros(1, device=dev), torch.zeros(1, device=dev)]
    down_stream = torch.cuda.Stream()          # the loss goes down on its own stream: the next step does not queue behind the copy
    e2e_state = {"i": 0, "losses": 0}

    def upload(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free_evt[b])          # the step that last read this half is done
            dbuf[b].copy_(host_data, non_blocking=True)
            up_evt[b].record(copy_stream)

    def step(resident=True):
        if resident:
            sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags,
                                     n_iter=1, step_sizes=steps_sz)
            return
        i = e2e_state["i"]
        b = i & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(up_evt[b])
        upload(i + 1)                                      # next step's volume, overlapped with this step
        sol.optimizing_transform(model=model, data=dbuf[b], init_output=init_out, optimize_flags=flags,
                                 n_iter=1, step_sizes=steps_sz)
        loss_dev[b].copy_(sol.last_dist.reshape(1))        # (the solver's own scalar is overwritten by the next step)
        free_evt[b].record(cur)
        with torch.cuda.stream(down_stream):
            down_stream.wait_event(free_evt[b])
            host_loss2[b].copy_(loss_dev[b], non_blocking=True)
            loss_evt[b].record(down_stream)
        if i > 0:                                          # read the previous step's loss on the host
            loss_evt[b ^ 1].synchronize()
            e2e_state["losses"] += 1 if float(host_loss2[b ^ 1][0]) == float(host_loss2[b ^ 1][0]) else 0
        e2e_state["i"] = i + 1

    def e2e_begin():
        e2e_state["i"] = 0
        free_evt[0].record(torch.cuda.current_stream())
        free_evt[1].record(torch.cuda.current_stream())
        upload(0)

    def e2e_end():
        i = e2e_state["i"]
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
            host_loss[0] = host_loss2[(i - 1) & 1][0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    REGIONS = 5

    def timed(n, resident, regions=1):
        """Times exactly n steps (barrier + synchronize on both sides, CUDA events, max over ranks).  With
        regions > 1 the n steps are cut into that many back-to-back sub-regions by extra events on the stream
        (no synchronisation in between). -> (total ms, [ms per step of every sub-region])"""
        barrier()
        regions = max(1, min(regions, n))
        cuts = [round(i * n / regions) for i in range(regions + 1)]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(regions + 1)]
        evs[0].record()
        if not resident:
            e2e_begin()
        r = 1
        for i in range(n):
            step(resident)
            if i + 1 == cuts[r] and r < regions:
                evs[r].record()
                r += 1
        if not resident:
            e2e_end()
        evs[regions].record()
        torch.cuda.synchronize()
        ms = evs[0].elapsed_time(evs[regions])
        per = [evs[i].elapsed_time(evs[i + 1]) / max(1, cuts[i + 1] - cuts[i]) for i in range(regions)]
        if world > 1:
            t = torch.tensor([ms] + per, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            vals = t.tolist()
            ms, per = float(vals[0]), [float(v) for v in vals[1:]]
        barrier()
        return ms, per

    # ---- eager warm-up; two steps run with every kernel bracketed by CUDA events -> per-kernel
    # breakdown and the dominant kernel
    eager_steps = min(args.steps, 50)          # the eager legs (breakdown / live roofline) are bounded
    sol.use_cuda_graph = False
    for _ in range(max(args.warmup, 3)):
        step(True)
    torch.cuda.synchronize()
    _lib.prof_configure("all", 16384)
    step(True)
    step(True)
    torch.cuda.synchronize()
    breakdown = _lib.prof_collect(16384)
    _lib.prof_configure(None)
    tot = {k: sum(v) / 2.0 for k, v in breakdown.items()}          # ms per step
    ranked = sorted(tot.items(), key=lambda kv: -kv[1])
    dom = next((k for k, _ in ranked if k in ALGO_WORDS), None)

    # ---- eager timed region: the dominant kernel bracketed by CUDA events on the launching stream
    # (events cannot bracket the nodes of a replayed graph, so the roofline is taken here)
    _lib.launch_count(reset=True)
    if dom is not None:
        _lib.prof_configure(dom, 16384)
    ms_eager, _ = timed(eager_steps, True)
    eager_launches = _lib.launch_count(reset=True) * args.steps // eager_steps
    dom_ms = _lib.prof_collect(16384).get(dom, []) if dom is not None else []
    _lib.prof_configure(None)
    # ... and once more with every RUN of back-to-back launches of that kernel bracketed by one event pair (the
    # 8 adjoint squaring steps of a field build): an event pair around each 40 us launch adds its own few
    # microseconds to every record; a run amortises them
    dom_runs = []
    if dom is not None:
        _lib.prof_group_runs(True)
        _lib.prof_configure(dom, 16384)
        timed(min(eager_steps, 20), True)
        dom_runs = _lib.prof_collect_runs(16384).get(dom, [])
        _lib.prof_configure(None)
        _lib.prof_group_runs(False)

    # ---- timed region 1 (`value`): inputs resident in HBM, PGD iteration replayed from a CUDA graph
    use_graph = not args.no_graph
    sol.use_cuda_graph = use_graph
    for _ in range(max(args.warmup, 3)):
        step(True)
    sol.graph_replays = 0
    if sampler:
        sampler.wait_ready()                  # started before the warm-up; make sure it is producing
    w0 = time.time()
    ms, region_ms = timed(args.steps, True, REGIONS)
    w1 = time.time()
    clocks = sampler.stop((w0, w1)) if sampler else None
    graph_used = use_graph and getattr(sol, "graph_replays", 0) == args.steps
    launches = (sol.graph_launches_per_replay * args.steps) if graph_used else eager_launches

    # ---- timed region 2: end to end through the public API with host buffers
    timed(2, False)
    ms_e2e, region_ms_e2e = timed(args.steps, False, REGIONS)

    if rank != 0:
        if world > 1:
            _release_graphs()
            dist.destroy_process_group()
        return 0

    nvox = _nvox(size)
    peak, peak_src = measured_peak_gbs()
    roof = None
    if dom_ms:
        avg_single = sum(dom_ms) / len(dom_ms)
        n_run = sum(c for _, c in dom_runs)
        avg_ms = (sum(t for t, _ in dom_runs) / n_run) if n_run else avg_single
        algo_bytes = ALGO_WORDS[dom](d) * 4.0 * nvox
        achieved = algo_bytes / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(dom) if args.workload == "m128" else None,   # captured on m128 only
                "peak_source": peak_src,
                "algo_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_ms,
                "launches_timed": n_run if n_run else len(dom_ms),
                "event_pairs": len(dom_runs) if n_run else len(dom_ms),
                "avg_launch_ms_one_event_pair_per_launch": avg_single,
                "frac_one_event_pair_per_launch": algo_bytes / (avg_single * 1e-3) / 1e9 / peak,
                "kernel_share_of_step": sum(dom_ms) / ms_eager if ms_eager > 0 else None,
                "measured_in": "eager timed regions (CUDA events on the launching stream; graph nodes cannot be "
                               "bracketed): %d steps with one event pair per launch (share of the step, "
                               "`*_one_event_pair_per_launch`), then %d steps with one event pair per run of "
                               "back-to-back launches of the kernel (`avg_launch_ms`, `achieved`, `frac`)"
                               % (eager_steps, min(eager_steps, 20))}
    # Roofline of the scopes SURVEY.md section 8d names: algorithmic words per voxel of the scope x 4 B x voxels
    # per GPU / the time our kernels of that scope take per step (event breakdown of the eager steps above).
    C, K = size[1], K_CLASSES
    has_int = "noise" in chain or "bias" in chain
    wA = (3 * C + d) if chain == ["noise", "bias", "morph", "affine"] else (2 * C + d)
    wD = (4 * C + 2 * d) if has_int else (2 * C + 2 * d)
    wB, wC_, wU = 2 * K + 3 * d + 1, 3 * K + 2 * d, (3 * C if "noise" in chain else 0)
    w_field = 94 * d
    # the same two chain passes as they can be EXECUTED: a resampling needs the complete output of the stage before
    # it, so every stage boundary is a round trip through HBM; compulsory words per stage
    has_noise, has_bias, has_morph, has_aff = ("noise" in chain), ("bias" in chain), ("morph" in chain), ("affine" in chain)
    wA_ps = ((2 + has_noise) * C if has_int else 0) + ((2 * C + d + 1) if has_morph else 0) + \
            ((2 * C + 1 + has_morph) if has_aff else 0)
    wD_ps = (3 * C if has_aff else 0) + ((3 * C + 2 * d) if has_morph else 0) + \
            (((2 + 2 * has_noise) * C + has_bias) if has_int else 0)
    words = wA + wB + wC_ + wD + wU + w_field        # full chain: 102d + 10C + 5K + 1
    FIELD_KERNELS = ("ss_step", "ss_step_bwd", "smooth_fwd", "smooth_bwd", "init_phi0", "lowres_smooth",
                     "adjoint_axis", "aos_to_planar", "unorm2", "steps_check")
    APPLY_KERNELS = ("chain_img_fwd", "chain_img_bwd", "chain_pk_fwd", "chain_pk_bwd", "chain_fwd", "chain_bwd",
                     "chain_fwd_stage", "chain_bwd_stage", "update", "sumsq", "lowfield_fwd", "lowfield_bwd",
                     "adjoint_axis_f", "affine_theta_fwd", "affine_theta_bwd")

    def scope(names, w):
        t = sum(tot.get(k, 0.0) for k in names)
        if t <= 0:
            return None
        gbs = w * 4.0 * nvox / (t * 1e-3) / 1e9
        return {"ms_per_step": round(t, 4), "algo_bytes_per_step": w * 4.0 * nvox, "achieved": gbs, "frac": gbs / peak,
                "kernels": [k for k in names if k in tot]}

    scopes = {
        "chain_apply_fwd_bwd_A+D": scope(("chain_img_fwd", "chain_img_bwd"), wA + wD),
        "chain_apply_fwd_bwd_A+D_per_stage": scope(("chain_img_fwd", "chain_img_bwd"), wA_ps + wD_ps),
        "apply_total_A+B+C+D+U": scope(APPLY_KERNELS, wA + wB + wC_ + wD + wU),
        "field_build_fwd_bwd": scope(FIELD_KERNELS, w_field),
        "whole_step_advk_kernels": {"ms_per_step": round(sum(tot.values()), 4), "algo_bytes_per_step": words * 4.0 * nvox,
                                    "achieved": words * 4.0 * nvox / (sum(tot.values()) * 1e-3) / 1e9,
                                    "frac": words * 4.0 * nvox / (sum(tot.values()) * 1e-3) / 1e9 / peak},
        "whole_step_incl_model_and_loss": {"ms_per_step": ms / args.steps, "algo_bytes_per_step": words * 4.0 * nvox,
                                           "achieved": words * 4.0 * nvox / (ms / args.steps * 1e-3) / 1e9,
                                           "frac": words * 4.0 * nvox / (ms / args.steps * 1e-3) / 1e9 / peak},
        "unit": "GB/s", "peak": peak,
        "note": "per-scope time = CUDA-event time of our kernels of that scope in eager steps (graph nodes cannot be "
                "bracketed); the last scope uses the CUDA-graph step time of `value`; `_per_stage` divides the same time "
                "by the compulsory bytes of the chain as executed (one HBM round trip per stage boundary: a resampling "
                "needs the complete output of the stage before it), the scope without the suffix by the bytes of a "
                "hypothetical single pass (SURVEY.md section 8d)",
    }
    iters = (1 if strong else world) * args.steps          # strong scaling: one iteration covers the global batch
    value = iters / (ms * 1e-3)
    med = statistics.median(region_ms)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, d, gsize, chain, per_gpu=size),
        "e2e": {"value": iters / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": host_data.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step_regions": [round(v, 4) for v in region_ms_e2e]},
        "gpu_launches": launches,
        "cuda_graph": bool(graph_used),
        "ms_per_step_regions": [round(v, 4) for v in region_ms],
        "ms_per_step_median_region": med,
        "value_median_region": (1 if strong else world) / (med * 1e-3),
        "ms_per_step_eager": ms_eager / eager_steps,
        "clocks": clocks,
        "roofline": roof,
        "roofline_scopes": scopes,
        "kernel_ms_per_step": {k: round(v, 4) for k, v in ranked},
        "advk_ms_per_step": round(sum(tot.values()), 4),
    }
    line["step_roofline"] = {
        "algo_bytes_per_step": words * 4.0 * nvox, "unit": "GB/s",
        "achieved_whole_step": words * 4.0 * nvox / (ms / args.steps * 1e-3) / 1e9,
        "achieved_advk_kernels_only": words * 4.0 * nvox / (sum(tot.values()) * 1e-3) / 1e9,
        "peak": peak}
    if not args.no_cuda_baseline and world == 1:
        # the incumbent: the unmodified reference, device=cuda, PyTorch eager, same GPU, same run
        try:
            ms_ref, ref_detail = time_reference_cuda(d, size, chain, dev, steps=5, warmup=3)
        except Exception as exc:                              # pragma: no cover
            ms_ref, ref_detail = None, None
            sys.stderr.write("bench.py: reference-on-CUDA leg failed: %s\n" % (exc,))
        line["cuda_eager_baseline"] = None if ms_ref is None else {
            "value": 1e3 / ms_ref, "unit": UNIT, "ms_per_step": ms_ref, "steps": 5, "warmup": 3,
            "ms_per_step_by_setting": ref_detail,
            "kind": "unmodified reference (baseline/_ref), device=cuda, PyTorch eager, same GPU; the faster of "
                    "torch.backends.cudnn.benchmark off / on",
            "speedup_resident": value * ms_ref / 1e3}
    else:
        line["cuda_eager_baseline"] = None
    if not args.no_cpu_baseline and world == 1:
        # bounded CPU sample at FULL size: one iteration of the CPU port (the real reference needs ~4x longer:
        # it builds every deformation field twice); `--impl reference` times the reference itself
        times, kind, cores = time_cpu_arm(d, size, chain, 1, 30.0, prefer_reference=False)
        line["cpu_baseline"] = {
            "value": len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d PGD inner-loop iteration(s) of the CPU port (oracle/) on %s at FULL size (no scaling), "
                      "%d host threads" % (len(times), "x".join(map(str, size)), cores)}
    else:
        line["cpu_baseline"] = None
    if args.profile_out:
        with open(args.profile_out, "w") as f:
            json.dump({"breakdown_ms_per_step": dict(ranked), "line": line}, f, indent=1)
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        _release_graphs()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-cuda":
        return run_reference_cuda(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
