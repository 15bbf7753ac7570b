"""GPU-box experiment: time the fused chain launches of one PGD step under tuning variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from advchain_b200 import _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "m128"
d, size, chain = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
torch.manual_seed(1234)
data = torch.rand(*size).to(dev)
torch.manual_seed(0)
conv = torch.nn.Conv2d if d == 2 else torch.nn.Conv3d
model = conv(size[1], bench.K_CLASSES, 3, 1, 1).eval().to(dev)
sol = bench.build_solver(d, size, chain, dev)
init_out = sol.get_init_output(model, data)
sol.init_random_transformation()
flags, steps = [True] * len(chain), [1.0] * len(chain)
lib = _lib.load()


def step():
    sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags,
                             n_iter=1, step_sizes=steps)


def measure(tag):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _lib.prof_configure("all", 16384)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    r = _lib.prof_collect(16384)
    _lib.prof_configure(None)
    out = []
    for k in ("chain_fwd", "chain_bwd", "chain_fwd_stage", "chain_bwd_stage"):
        if k in r:
            v = r[k]
            per = len(v) // 3
            out.append("%s: %s" % (k, " ".join("%.0f" % (1e3 * sum(v[i::per]) / 3) for i in range(per))))
    print("%-28s %s" % (tag, " | ".join(out)), flush=True)


for pack in (1, 0):
    lib.advk_chain_set_packed(pack)
    for coop in (1, 0):
        lib.advk_chain_set_cooperative(coop)
        for minb in ((43, 42, 44, 33) if coop else (43,)):
            lib.advk_chain_tune(minb, 1)
            measure("pack=%d coop=%d minb(fwd,bwd)=%d" % (pack, coop, minb))
