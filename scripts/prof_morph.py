"""Profiling driver for the morph field build: warm-up, then ONE forward + backward build per kernel
variant between cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).

    python scripts/prof_morph.py [workload] [mask ...]
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
masks = [int(a) for a in sys.argv[2:]] or [0, 1]
d, size, chain = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
cfg = bench.make_cfgs(d, size)["morph"]
lib = _lib.load()
t = AdvMorph(d, cfg, device=dev)
t.init_parameters()
v = t.param.detach().clone()
gout = None


def build():
    global gout
    t.param = v.clone().requires_grad_(True)
    t._cache.clear()
    f = t._field(1)
    if gout is None:
        torch.manual_seed(3)
        gout = torch.randn_like(f)
    f.backward(gout)


for m in masks:
    lib.advk_morph_tune(m)
    build()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for m in masks:
    lib.advk_morph_tune(m)
    build()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
