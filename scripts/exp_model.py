"""How the toy model of the bench (Conv3d(1,4,3,1,1) on 1x1x128^3: forward + input gradient, what one PGD iteration
asks of it) fares under the PyTorch backend flags a user can set.  python scripts/exp_model.py"""
import itertools
import torch

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = torch.nn.Conv3d(1, 4, 3, 1, 1).eval().to(dev)
x = torch.rand(1, 1, 128, 128, 128, device=dev, requires_grad=True)
g = torch.rand(1, 4, 128, 128, 128, device=dev)


def run(n):
    for _ in range(n):
        y = model(x)
        (gx,) = torch.autograd.grad(y, x, g)
    return gx


ref = None
for enabled, bench, tf32 in itertools.product((True, False), (True, False), (True, False)):
    if not enabled and (bench or not tf32):
        continue
    torch.backends.cudnn.enabled = enabled
    torch.backends.cudnn.benchmark = bench
    torch.backends.cudnn.allow_tf32 = tf32
    gx = run(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(20)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = gx.clone()
    print("cudnn.enabled=%s benchmark=%s allow_tf32=%s: %.1f us per fwd+dgrad, max |dgrad - first| %.2e"
          % (enabled, bench, tf32, 1e3 * e0.elapsed_time(e1) / 20, float((gx - ref).abs().max())))
