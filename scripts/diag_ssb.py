"""GPU-box diagnostic: where does a launch of the squaring-step adjoint spend its time?

    python scripts/diag_ssb.py <0..4> [workload]

Loads scripts/_diag/libadvk_diag<k>.so (scripts/build_diag.sh; k = 0: the product library) in place of
the product library and times `ss_step_bwd` inside the morph field build.  The side libraries compute
WRONG gradients by construction (REDs replaced by register adds / plain stores, gathers replaced by the
thread's own value): only the launch times mean anything.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from advchain_b200 import _lib  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 0
wl = sys.argv[2] if len(sys.argv) > 2 else "m128"
if k:
    _lib.LIB_PATH = os.path.join(ROOT, "scripts", "_diag", "libadvk_diag%d.so" % k)
from advchain_b200.augmentor import AdvMorph  # noqa: E402

d, size, chain = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
cfg = bench.make_cfgs(d, size)["morph"]
t = AdvMorph(d, cfg, device=dev)
t.init_parameters()
v_unit = t.param.detach().clone()
reps = 5
for vn in (1.0, 4.0):
    t.param = (v_unit * vn).clone().requires_grad_(True)
    gout = None
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
    torch.cuda.synchronize()
    r = _lib.prof_collect(16384)
    _lib.prof_configure(None)
    ssb = r.get("ss_step_bwd", [])
    print("diag %d vnorm %.1f: ss_step_bwd %.1f us per launch (%d launches), ss_step %.1f us per launch"
          % (k, vn, 1e3 * sum(ssb) / max(len(ssb), 1), len(ssb),
             1e3 * sum(r.get("ss_step", [0])) / max(len(r.get("ss_step", [0])), 1)), flush=True)
