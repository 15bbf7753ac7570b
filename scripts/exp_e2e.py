"""GPU-box experiment: where do the ~0.1 ms per step go that separate the end-to-end leg of bench.py from the
resident one?  Times the graph-replayed PGD step under combinations of {per-step H2D upload, per-step loss
download + host wait with a lag of L steps}."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

d, size, chain = bench.WORKLOADS["m128"]
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234)
host = torch.rand(*size).pin_memory()
data = host.to(dev)
torch.manual_seed(0)
model = torch.nn.Conv3d(size[1], bench.K_CLASSES, 3, 1, 1).eval().to(dev)
sol = bench.build_solver(d, size, chain, dev)
init_out = sol.get_init_output(model, data)
sol.init_random_transformation()
flags, steps = [True] * len(chain), [1.0] * len(chain)
sol.use_cuda_graph = True
copy_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
NB = 4
dbuf = [torch.empty_like(data) for _ in range(NB)]
up_evt = [torch.cuda.Event() for _ in range(NB)]
free_evt = [torch.cuda.Event() for _ in range(NB)]
loss_dev = [torch.zeros(1, device=dev) for _ in range(NB)]
loss_host = [torch.zeros(1).pin_memory() for _ in range(NB)]
loss_evt = [torch.cuda.Event() for _ in range(NB)]


def run(n, upload, download, lag, own_stream=True):
    cur = torch.cuda.current_stream()
    for e in free_evt:
        e.record(cur)
    if upload:
        with torch.cuda.stream(copy_stream):
            dbuf[0].copy_(host, non_blocking=True)
            up_evt[0].record(copy_stream)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    host_wait = 0.0
    t0.record()
    for i in range(n):
        b = i % NB
        x = data
        if upload:
            cur.wait_event(up_evt[b])
            nb_ = (i + 1) % NB
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free_evt[nb_])
                dbuf[nb_].copy_(host, non_blocking=True)
                up_evt[nb_].record(copy_stream)
            x = dbuf[b]
        sol.optimizing_transform(model=model, data=x, init_output=init_out, optimize_flags=flags, n_iter=1, step_sizes=steps)
        if download:
            loss_dev[b].copy_(sol.last_dist.reshape(1))
        free_evt[b].record(cur)
        if download:
            if own_stream:
                with torch.cuda.stream(down_stream):
                    down_stream.wait_event(free_evt[b])
                    loss_host[b].copy_(loss_dev[b], non_blocking=True)
                    loss_evt[b].record(down_stream)
            else:
                loss_host[b].copy_(loss_dev[b], non_blocking=True)
                loss_evt[b].record(cur)
            if i >= lag:
                w0 = time.perf_counter()
                loss_evt[(i - lag) % NB].synchronize()
                host_wait += time.perf_counter() - w0
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n, 1e3 * host_wait / n


for _ in range(5):
    sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags, n_iter=1, step_sizes=steps)
torch.cuda.synchronize()
# host time of one step's Python + enqueue (GPU queue never empty: no sync)
w0 = time.perf_counter()
for _ in range(50):
    sol.optimizing_transform(model=model, data=data, init_output=init_out, optimize_flags=flags, n_iter=1, step_sizes=steps)
w1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue time per step (resident, queue running ahead): %.3f ms" % (1e3 * (w1 - w0) / 50), flush=True)
for name, kw in [("resident", dict(upload=False, download=False, lag=1)),
                 ("upload only", dict(upload=True, download=False, lag=1)),
                 ("download lag 1 (own stream)", dict(upload=False, download=True, lag=1)),
                 ("download lag 1 (compute stream)", dict(upload=False, download=True, lag=1, own_stream=False)),
                 ("download lag 2", dict(upload=False, download=True, lag=2)),
                 ("download lag 3", dict(upload=False, download=True, lag=3)),
                 ("upload + download lag 1", dict(upload=True, download=True, lag=1)),
                 ("upload + download lag 2", dict(upload=True, download=True, lag=2)),
                 ("resident again", dict(upload=False, download=False, lag=1))]:
    run(10, **kw)
    ms, hw = run(100, **kw)
    print("%-34s %.4f ms / step   host blocked %.3f ms / step" % (name, ms, hw), flush=True)
