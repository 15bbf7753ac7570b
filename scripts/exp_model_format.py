"""GPU-box experiment: cost of the benchmark's toy model Conv3d(1,4,3,1,1) (forward + input gradient)
in the default and the channels_last_3d memory format, as cuDNN picks its kernels."""
import torch
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
N, S = 1, 128


def run(fmt, tf32=True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.manual_seed(0)
    m = torch.nn.Conv3d(1, 4, 3, 1, 1).eval().to(dev)
    x = torch.rand(N, 1, S, S, S, device=dev)
    go = torch.randn(N, 4, S, S, S, device=dev)
    if fmt == "cl":
        m = m.to(memory_format=torch.channels_last_3d)
        x = x.contiguous(memory_format=torch.channels_last_3d)
        go = go.contiguous(memory_format=torch.channels_last_3d)
    x.requires_grad_(True)

    def it():
        y = m(x)
        (gx,) = torch.autograd.grad(y, x, go)
        return y, gx
    for _ in range(5):
        y, gx = it()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y, gx = it()
    e1.record()
    torch.cuda.synchronize()
    print("%s tf32=%s: %.1f us per fwd+dgrad; y strides %s cl=%s; gx strides %s" % (
        fmt, tf32, 1e3 * e0.elapsed_time(e1) / 20, tuple(y.stride()),
        y.is_contiguous(memory_format=torch.channels_last_3d), tuple(gx.stride())), flush=True)
    return y.detach(), gx.detach()


a = run("nchw")
b = run("cl")
print("max diff y %.3e gx %.3e" % ((a[0] - b[0]).abs().max().item(), (a[1] - b[1]).abs().max().item()))
run("nchw", tf32=False)
run("cl", tf32=False)
