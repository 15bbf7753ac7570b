"""GPU-box experiment: what the toy model Conv3d(1,4,3,1,1) (fwd + input-gradient) costs under cuDNN settings."""
import itertools
import torch

dev = torch.device("cuda:0")
x = torch.rand(1, 1, 128, 128, 128, device=dev)
for tf32, bench in itertools.product([True, False], [True, False]):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    m = torch.nn.Conv3d(1, 4, 3, 1, 1).eval().to(dev)
    xx = x.clone().requires_grad_(True)
    g = torch.randn(1, 4, 128, 128, 128, device=dev)

    def step():
        y = m(xx)
        (gx,) = torch.autograd.grad(y, xx, g)
        return gx

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("allow_tf32=%s benchmark=%s: %.1f us per fwd+dgrad" % (tf32, bench, 1e3 * e0.elapsed_time(e1) / 20))
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
        print("     %8.1f us  %s" % (ev.device_time_total, ev.key[:90]))
