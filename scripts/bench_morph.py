"""GPU-box experiment: time the kernels of the morph field build (forward + backward) under the
squaring-step kernel variants (advk_morph_tune) and check that the variants agree.

    python scripts/bench_morph.py [workload] [vnorm ...]

Environment: BENCH_MORPH_MASKS = comma-separated advk_morph_tune masks to time (default: all variants),
BENCH_MORPH_OUT = JSON file that receives {mask: us of ss_step_bwd per build} summed over the vnorms
plus "best" (fastest mask whose field and gradient agree with mask 0).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from advchain_b200 import _lib  # noqa: E402
from advchain_b200.augmentor import AdvMorph  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "m128"
vnorms = [float(a) for a in sys.argv[2:]] or [1.0, 4.0]
d, size, chain = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
cfg = bench.make_cfgs(d, size)["morph"]
lib = _lib.load()
t = AdvMorph(d, cfg, device=dev)
t.init_parameters()
torch.manual_seed(3)
v_unit = t.param.detach().clone()
gout = None


def run(vn, mask, reps=5):
    global gout
    lib.advk_morph_tune(mask)
    t.param = (v_unit * vn).clone().requires_grad_(True)
    res = None
    for it in range(reps + 2):
        if it == 2:
            torch.cuda.synchronize()
            _lib.prof_configure("all", 16384)
        t.param.grad = None
        t._cache.clear()
        f = t._field(1)
        if gout is None:
            gout = torch.randn_like(f)
        f.backward(gout)
        res = (f.detach().clone(), t.param.grad.detach().clone())
    torch.cuda.synchronize()
    r = _lib.prof_collect(16384)
    _lib.prof_configure(None)
    line = " ".join("%s=%.1f" % (k, 1e3 * sum(vs) / reps) for k, vs in sorted(r.items(), key=lambda kv: -sum(kv[1])))
    print("vnorm %.1f tile-mask %d  [us per build, %d reps]: %s" % (vn, mask, reps, line), flush=True)
    ssb_us[mask] = ssb_us.get(mask, 0.0) + 1e3 * sum(r.get("ss_step_bwd", [0.0])) / reps
    return res


ssb_us = {}
agree = {}
masks = [int(m) for m in os.environ.get("BENCH_MORPH_MASKS", "0").split(",")]
for vn in vnorms:
    base = run(vn, 1)
    for mask in masks:
        out = run(vn, mask)
        ef = float((out[0] - base[0]).abs().max() / base[0].abs().max())
        eg = float((out[1] - base[1]).abs().max() / base[1].abs().max())
        agree[mask] = agree.get(mask, True) and ef < 1e-6 and eg < 1e-4
        print("   mask %d vs 1 (plain): field rel err %.2e, grad rel err %.2e" % (mask, ef, eg), flush=True)
ok = [m for m in masks if agree.get(m)]
best = min(ok, key=lambda m: ssb_us[m]) if ok else 0
print("ss_step_bwd us per build (sum over vnorms): %s -> best %d" % (
    " ".join("%d=%.1f" % (m, ssb_us[m]) for m in [1] + masks), best), flush=True)
if os.environ.get("BENCH_MORPH_OUT"):
    with open(os.environ["BENCH_MORPH_OUT"], "w") as fh:
        json.dump({"workload": wl, "us": {str(m): ssb_us[m] for m in ssb_us}, "agree": {str(m): bool(v) for m, v in agree.items()},
                   "best": best}, fh)
