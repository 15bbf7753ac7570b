"""Shared test utilities: fixture loading and oracle construction."""
import json
import os

import numpy as np
import torch

from oracle import advchain_oracle as orc
from tests.golden.cases import CASES, stage_cfgs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, arrays


def rel_err(a, b):
    """max|a-b| / max|b| -- the normwise metric of SURVEY.md section 8c."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def make_model(case, arrays, device="cpu"):
    conv = torch.nn.Conv2d if case["d"] == 2 else torch.nn.Conv3d
    m = conv(case["size"][1], case["K"], 3, 1, 1)
    with torch.no_grad():
        m.weight.copy_(arrays["model_w"])
        m.bias.copy_(arrays["model_b"])
    return m.eval().to(device)


def oracle_solver(case, size=None):
    size = size or case["size"]
    cfgs = stage_cfgs(case["d"], size, vector=case.get("vector"))
    pad = case.get("padding", "zeros")
    stages = []
    for n in case["chain"]:
        kw = {"padding": pad} if n in ("morph", "affine") else {}
        st = orc.make_stage(n, cfgs[n], **kw)
        st.power_iteration = bool(case.get("power", False))
        stages.append(st)
    return orc.Solver(stages, if_norm_image=True)


def cuda_solver(case, device, size=None, **solver_kw):
    """The product (CUDA) transforms + solver for a golden case."""
    from advchain_b200.augmentor import (AdvAffine, AdvBias, AdvMorph, AdvNoise,
                                         ComposeAdversarialTransformSolver)
    size = size or case["size"]
    d = case["d"]
    cfgs = stage_cfgs(d, size, vector=case.get("vector"))
    pad = case.get("padding", "zeros")
    chain = []
    pw = bool(case.get("power", False))
    for n in case["chain"]:
        if n == "noise":
            chain.append(AdvNoise(d, cfgs[n], power_iteration=pw, device=device))
        elif n == "bias":
            chain.append(AdvBias(d, cfgs[n], power_iteration=pw, device=device))
        elif n == "morph":
            chain.append(AdvMorph(d, cfgs[n], power_iteration=pw, image_padding_mode=pad, device=device))
        else:
            chain.append(AdvAffine(d, cfgs[n], power_iteration=pw, image_padding_mode=pad, device=device))
    kw = dict(divergence_types=["mse", "contour"], divergence_weights=[1.0, 0.5], if_norm_image=True)
    kw.update(solver_kw)
    return ComposeAdversarialTransformSolver(chain, **kw)


def oracle_replay(case, z, params, dtype=torch.float64, train=True, size=None):
    """Teacher-forced replay of one PGD step (or, with train=False, of the final eval-mode chain)
    by the CPU oracle in `dtype`.  With float64 it gives the rounding-free value of every quantity
    the fixtures hold in fp32, i.e. the reference's own fp32 noise floor
    floor(q) = rel_err(fixture_q, oracle64_q)."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        sol = oracle_solver(case, size)
        model = make_model(case, z).to(dtype)
        data, init_out = z["data"].to(dtype), z["init_output"].to(dtype)
        for s, p in zip(sol.stages, params):
            s.init()
            s.param = p.to(dtype)
        out = {}
        if train:
            for s in sol.stages:
                s.train()
            dist, pred, mask = sol.step_loss(model, data, init_out)
            dist.backward()
            out["dist"], out["pred"] = dist.detach(), pred.detach()
            out["grads"] = [s.param.grad.detach() for s in sol.stages]
        else:
            with torch.no_grad():
                out["adv"] = sol.forward(data)
                out["pf"] = sol.predict_forward(init_out)
                out["pb"] = sol.predict_backward(z["logits"].to(dtype))
        return out
    finally:
        torch.set_default_dtype(old)


def noise_floor(fixture_value, value64):
    """The reference's own fp32 rounding error for a quantity, measured against the fp64 oracle."""
    return rel_err(fixture_value, value64)
