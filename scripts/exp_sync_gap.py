"""GPU-box experiment: how much of a graph-mode PGD iteration is the per-call host synchronisation?
Compares 10 calls of optimizing_transform(n_iter=1) with 1 call of optimizing_transform(n_iter=10)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

d, size, chain = bench.WORKLOADS["m128"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
data = torch.rand(*size, device=dev)
model = torch.nn.Conv3d(1, 4, 3, 1, 1).eval().to(dev)
sol = bench.build_solver(d, size, chain, dev)
sol.use_cuda_graph = True
init = sol.get_init_output(model, data)
sol.init_random_transformation()
flags, steps = [True] * 4, [1.0] * 4


def run(n_iter, calls):
    for _ in range(3):
        sol.optimizing_transform(model=model, data=data, init_output=init, optimize_flags=flags, n_iter=n_iter, step_sizes=steps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        sol.optimizing_transform(model=model, data=data, init_output=init, optimize_flags=flags, n_iter=n_iter, step_sizes=steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (calls * n_iter)


a = run(1, 100)
b = run(10, 10)
print("ms per PGD iteration: n_iter=1 per call %.4f ; n_iter=10 per call %.4f ; per-call overhead ~%.0f us"
      % (a, b, 1e3 * (a - b) * 10 / 9))
