"""GPU-box experiment: a multi-iteration PGD call (BASELINE.json configs 2-5 run 5-10 inner steps per call) through
the CUDA-graph loop; ms per inner iteration, with the speculative 3-D step count (default) or the exact norm read
behind every replay (ADVK_SPECULATE=0).    python scripts/exp_multi.py [workload] [n_iter] [calls]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "m128"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 5
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 20
d, size, chain = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234)
data = torch.rand(*size).to(dev)
conv = torch.nn.Conv2d if d == 2 else torch.nn.Conv3d
torch.manual_seed(0)
model = conv(size[1], bench.K_CLASSES, 3, 1, 1).eval().to(dev)
sol = bench.build_solver(d, size, chain, dev)
init_out = sol.get_init_output(model, data)
flags, steps = [True] * len(chain), [1.0] * len(chain)
sol.use_cuda_graph = True
torch.manual_seed(7)
sol.init_random_transformation()
start = [t.param.detach().clone() for t in sol.chain_of_transforms]


def call():
    for t, p in zip(sol.chain_of_transforms, start):      # every call starts from the same parameters
        t.param = p.clone()
    sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags, n_iter=n_iter,
                             step_sizes=steps)


for _ in range(4):
    call()
torch.cuda.synchronize()
r0 = getattr(sol, "graph_replays", 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(calls):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("%s n_iter %d ADVK_SPECULATE=%s: %.4f ms per inner iteration (%.1f it/s), %d replays for %d iterations, "
      "%d iteration redos, %d loop redos" % (
          wl, n_iter, os.environ.get("ADVK_SPECULATE", "1"), ms / (calls * n_iter), 1e3 * calls * n_iter / ms,
          getattr(sol, "graph_replays", 0) - r0, calls * n_iter, getattr(sol, "graph_iter_redos", 0),
          getattr(sol, "graph_redos", 0)))
print("   last call (iteration, |u| of the previous one, band, count, predicted): %s" % (getattr(sol, "spec_trace", None),))
