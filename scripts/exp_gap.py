"""Host time between the end-of-call synchronisation of one graph-mode PGD call and the replay of the next one
(the GPU is idle for it): where the per-call host overhead goes.  python scripts/exp_gap.py"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

d, size, chain = bench.WORKLOADS["m128"]
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234)
data = torch.rand(*size).to(dev)
torch.manual_seed(0)
model = torch.nn.Conv3d(size[1], bench.K_CLASSES, 3, 1, 1).eval().to(dev)
sol = bench.build_solver(d, size, chain, dev)
init_out = sol.get_init_output(model, data)
sol.init_random_transformation()
flags, steps = [True] * len(chain), [1.0] * len(chain)
sol.use_cuda_graph = True
marks = {"verified": [], "replay_in": [], "replay_out": []}
orig_replay = torch.cuda.CUDAGraph.replay
orig_verified = sol._graph_verified


def replay(self):
    marks["replay_in"].append(time.perf_counter())
    orig_replay(self)
    marks["replay_out"].append(time.perf_counter())


def verified():
    r = orig_verified()
    marks["verified"].append(time.perf_counter())
    return r


torch.cuda.CUDAGraph.replay = replay
sol._graph_verified = verified


def step():
    sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags, n_iter=1,
                             step_sizes=steps)


for _ in range(5):
    step()
torch.cuda.synchronize()
for k in marks:
    marks[k].clear()
t0 = time.perf_counter()
N = 100
for _ in range(N):
    step()
torch.cuda.synchronize()
t1 = time.perf_counter()
gap = [1e6 * (marks["replay_in"][i + 1] - marks["verified"][i]) for i in range(N - 1)]
launch = [1e6 * (b - a) for a, b in zip(marks["replay_in"], marks["replay_out"])]
wait = [1e6 * (marks["verified"][i] - marks["replay_out"][i]) for i in range(N)]
print("per call %.1f us | sync return -> next replay() entered: median %.1f us | inside replay(): %.1f us | replay() "
      "returned -> sync returned (finish-loop host work + waiting for the GPU): %.1f us"
      % (1e6 * (t1 - t0) / N, statistics.median(gap), statistics.median(launch), statistics.median(wait)))
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
