"""GPU-box experiment: time the kernels of the morph field build (forward + backward) under the
squaring-step kernel variants (advk_morph_tune) and check that the variants agree.

    python scripts/bench_morph.py [workload] [vnorm ...]
"""
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
    return res


for vn in vnorms:
    base = run(vn, 0)
    for mask in (1, 8, 9):
        out = run(vn, mask)
        ef = float((out[0] - base[0]).abs().max() / base[0].abs().max())
        eg = float((out[1] - base[1]).abs().max() / base[1].abs().max())
        print("   mask %d vs 0: field rel err %.2e, grad rel err %.2e" % (mask, ef, eg), flush=True)
