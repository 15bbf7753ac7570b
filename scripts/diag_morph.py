"""GPU-box diagnostic for the morph field build backward: error matrix over geometries."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import advchain_oracle as orc
from tests.golden.cases import stage_cfgs
from tests.helpers import rel_err
from advchain_b200.augmentor import AdvMorph

dev = torch.device("cuda:0")


def fcf(field, d):
    return field[..., :d].permute(0, d + 1, *range(1, d + 1)).contiguous()


def run(d, size, vsize, scale, vnorm, seed=4, gmode="randn"):
    torch.manual_seed(seed)
    n = size[0]
    cfg = stage_cfgs(d, size, vector=vsize)["morph"]
    t = AdvMorph(d, cfg, device=dev)
    v = orc.unit_l2(torch.rand(n, d, *vsize) * 2 - 1) * vnorm
    gout = torch.randn(n, d, *size[2:])
    if gmode == "interior":      # no gradient near the faces: isolates boundary handling
        m = torch.zeros_like(gout)
        sl = (slice(None), slice(None)) + tuple(slice(3, s - 3) for s in size[2:])
        m[sl] = 1
        gout = gout * m
    res = {}
    for dt in (torch.float32, torch.float64):
        torch.set_default_dtype(dt)
        v0 = v.to(dt).clone().requires_grad_(True)
        ref = orc.morph_field(v0, scale, size[2:])
        ref.backward(gout.to(dt))
        res[dt] = (ref.detach(), v0.grad.detach())
    torch.set_default_dtype(torch.float32)
    t.param = v.to(dev).requires_grad_(True)
    t.epsilon = abs(scale)
    field = t._field(1 if scale > 0 else -1)
    out = torch.clamp(fcf(field, d), -1, 1)
    out.backward(gout.to(dev))
    g = t.param.grad
    print("%dD %-18s v%-10s s%+.1f |v|=%g %-8s out %.1e | grad cuda~32 %.2e cuda~64 %.2e 32~64 %.2e" % (
        d, size, vsize, scale, vnorm, gmode, rel_err(out, res[torch.float32][0]),
        rel_err(g, res[torch.float32][1]), rel_err(g, res[torch.float64][1]),
        rel_err(res[torch.float32][1], res[torch.float64][1])))
    return g.cpu(), res[torch.float32][1]


for gm in ("randn", "interior"):
    run(3, [2, 1, 16, 24, 40], [2, 3, 4], 1.5, 1.0, gmode=gm)
    run(3, [1, 1, 16, 24, 40], [2, 3, 4], 1.5, 1.0, gmode=gm)
    run(3, [2, 1, 16, 24, 40], [3, 3, 4], 1.5, 1.0, gmode=gm)
    run(3, [2, 1, 16, 24, 40], [4, 4, 4], 1.5, 1.0, gmode=gm)
    run(3, [2, 1, 16, 24, 40], [2, 3, 4], 1.5, 6.0, gmode=gm)
    run(3, [2, 1, 16, 24, 40], [2, 3, 4], -1.5, 1.0, gmode=gm)
    run(3, [2, 1, 32, 32, 32], [2, 2, 2], 1.5, 1.0, gmode=gm)
    run(3, [1, 1, 20, 20, 12], [3, 3, 2], -1.5, 1.0, gmode=gm)
    run(2, [2, 1, 48, 80], [3, 5], 1.5, 1.0, gmode=gm)
for seed in range(5, 9):
    run(3, [2, 1, 16, 24, 40], [2, 3, 4], 1.5, 1.0, seed=seed)
g, r = run(3, [2, 1, 16, 24, 40], [2, 3, 4], 1.5, 1.0)
e = (g - r).abs()
print("max err at", (e == e.max()).nonzero().tolist(), "err", e.max().item(), "ref max", r.abs().max().item())
print("err per (n,ch):", e.amax(dim=(2, 3, 4)))
print("err along D:", e.amax(dim=(0, 1, 3, 4)), "along H:", e.amax(dim=(0, 1, 2, 4)), "along W:", e.amax(dim=(0, 1, 2, 3)))
