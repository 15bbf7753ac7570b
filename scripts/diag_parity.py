"""GPU-box diagnostic: where do the CUDA path and the CPU oracle differ, and how does that compare
with the oracle's own fp32-vs-fp64 rounding noise?  (Test infrastructure; uses oracle/.)

    python scripts/diag_parity.py [case] [step]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tests.helpers import cuda_solver, load_golden, make_model, oracle_solver, rel_err  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2d_full"
step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
meta, z = load_golden(name)
case = meta["case"]
d = case["d"]
sp = case["size"][2:]


def oracle_run(dtype):
    torch.set_default_dtype(dtype)
    osol = oracle_solver(case)
    model = make_model(case, z).to(dtype)
    data, init_out = z["data"].to(dtype), z["init_output"].to(dtype)
    for i, s in enumerate(osol.stages):
        s.init()
        s.param = z["s%d_param_%d" % (step, i)].to(dtype)
        s.train()
    inter = {}
    t = data.clone()
    for s in osol.stages:
        t = s.fwd(t)
        inter["after_" + s.name] = t.detach()
        if s.name == "morph":
            inter["phi+"] = s.field(+1).detach()
            inter["phi-"] = s.field(-1).detach()
    dist, pred, mask = osol.step_loss(model, data, init_out)
    dist.backward()
    inter["pred"] = pred.detach()
    inter["dist"] = dist.detach()
    for s in osol.stages:
        inter["grad_" + s.name] = s.param.grad.detach()
    torch.set_default_dtype(torch.float32)
    return inter


o32, o64 = oracle_run(torch.float32), oracle_run(torch.float64)

sol = cuda_solver(case, dev)
model = make_model(case, z, dev)
data, init_out = z["data"].to(dev), z["init_output"].to(dev)
chain = sol.chain_of_transforms
mine = {}
for i, t in enumerate(chain):
    t.init_parameters()
    t.param = z["s%d_param_%d" % (step, i)].to(dev)
    t.train()
t_data = data
for t in chain:
    t_data = t.forward(t_data)
    mine["after_" + t.get_name()] = t_data.detach()
    if t.get_name() == "morph":
        for sign, key in ((+1, "phi+"), (-1, "phi-")):
            f = torch.clamp(t._field(sign).detach()[..., :d], -1, 1)
            mine[key] = f.permute(0, d + 1, *range(1, d + 1)).contiguous()
aug = sol.forward(data)
pred = sol.predict_backward(model(aug))
dist = sol.loss_fn(pred, init_out, sol.valid_region_mask(init_out))
dist.backward()
mine["pred"] = pred.detach()
mine["dist"] = dist.detach()
for t in chain:
    mine["grad_" + t.get_name()] = t.param.grad.detach()

print("case %s step %d   (rel = max|a-b|/max|b|)" % (name, step))
print("%-14s %12s %12s %12s" % ("quantity", "cuda~ref32", "cuda~ref64", "ref32~ref64"))
for k in o32:
    if k not in mine:
        continue
    a = mine[k].cpu()
    print("%-14s %12.3e %12.3e %12.3e" % (k, rel_err(a, o32[k]), rel_err(a, o64[k]), rel_err(o32[k], o64[k])))
for key in ("phi+", "phi-"):
    if key in mine:
        for ax in range(d):
            size = sp[d - 1 - ax]
            e = lambda a, b: ((a[:, ax].double().cpu() - b[:, ax].double()).abs().max().item() * (size - 1) / 2)
            print("%s axis %d error in px: cuda~ref32 %.3e cuda~ref64 %.3e ref32~ref64 %.3e" % (
                key, ax, e(mine[key], o32[key]), e(mine[key], o64[key]), e(o32[key], o64[key])))
