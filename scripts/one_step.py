"""Profiling driver: N warm-up PGD inner-loop iterations, then ONE iteration between
cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).

    python scripts/one_step.py [--workload m128] [--warmup 2] [--steps 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="m128")
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
d, size, chain = bench.WORKLOADS[a.workload]
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


def step():
    sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags,
                             n_iter=1, step_sizes=steps)


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done, last dist", float(sol.last_dist))
