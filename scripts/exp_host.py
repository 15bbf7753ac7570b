"""GPU-box experiment: host-side cost of one graph-replayed optimizing_transform call, by segment."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

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


def step():
    sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags, n_iter=1, step_sizes=steps)


for _ in range(5):
    step()
torch.cuda.synchronize()
# 1. GPU time of the replay alone (graph launched back to back)
key = [k for k, v in sol._graphs.items() if isinstance(v, dict)][0]
st = sol._graphs[key]
g = list(st["graphs"].values())[0][0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("graph replay back to back: %.4f ms" % (e0.elapsed_time(e1) / 50), flush=True)
# 2. the public call, wall clock with the GPU drained before every call: host prep + GPU + tail
ts = []
for _ in range(30):
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    step()
    w1 = time.perf_counter()
    torch.cuda.synchronize()
    w2 = time.perf_counter()
    ts.append((w1 - w0, w2 - w1))
print("call returns after %.3f ms (median), trailing GPU work %.3f ms" % (1e3 * sorted(t[0] for t in ts)[15], 1e3 * sorted(t[1] for t in ts)[15]), flush=True)
w0 = time.perf_counter()
for _ in range(100):
    step()
torch.cuda.synchronize()
print("100 calls back to back: %.4f ms / call" % (1e3 * (time.perf_counter() - w0) / 100), flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
