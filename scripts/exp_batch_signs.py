"""GPU-box experiment: does building both signs of the deformation field in ONE batched launch
sequence (2N samples) beat two sequences of N samples?  Times forward + backward of the field build,
events around the whole sequence, CUDA-graph replayed to take the host out of the picture."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from advchain_b200.augmentor import AdvMorph  # noqa: E402

dev = torch.device("cuda:0")


def make(n):
    size = [n, 1, 128, 128, 128]
    cfg = bench.make_cfgs(3, size)["morph"]
    t = AdvMorph(3, cfg, device=dev)
    t.init_parameters()
    return t


def build(t, gout, signs):
    t.param = t.param.detach().clone().requires_grad_(True)
    t._cache.clear()
    for s in signs:
        f = t._field(s)
        f.backward(gout)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


t1 = make(1)
t1._fixed_steps = None
g1 = torch.randn(1, 128, 128, 128, 4, device=dev)
t2 = make(2)
g2 = torch.randn(2, 128, 128, 128, 4, device=dev)
# pin the step count (no host read inside the captured region)
for t in (t1, t2):
    n = t._nb_steps()
    viol = torch.zeros(1, dtype=torch.int32, device=dev)
    t._fixed_steps = (n, viol)
a = timed(lambda: build(t1, g1, (1, -1)))
b = timed(lambda: build(t2, g2, (1,)))
print("two sequences of N=1 (both signs): %.1f us ; one sequence of N=2: %.1f us" % (a, b))
