"""fp32-vs-fp64 noise floor of the reference ALGORITHM at the full-size cases (build container or anywhere:
needs only the oracle, which tests/test_oracle_golden_big.py pins to the reference at these sizes).

    python tests/golden/make_floors_big.py [case ...]   ->  tests/golden/<case>.floors.json

For every quantity the GPU parity test compares (dist, warped-back prediction, raw parameter gradients of
each replayed step; final chain outputs) it records rel_err(fp32 evaluation, fp64 evaluation) of the SAME
algorithm from the SAME start parameters -- how far two legitimate floating-point evaluations of the
reference lie apart on white-noise volumes.  The GPU test allows max(tolerance, 2 x floor)
(tests/test_gpu_golden.py explains the factor)."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.bighelpers import load_big, start_params  # noqa: E402
from tests.golden.cases import BIG_CASES  # noqa: E402
from tests.helpers import make_model, oracle_solver, rel_err  # noqa: E402


def step(sol, model, data, init_out, params, dtype):
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        for st, p in zip(sol.stages, params):
            st.param = p.to(dtype).clone()
            st.train()
        m = model.to(dtype)
        m.zero_grad()
        dist, pred, _ = sol.step_loss(m, data.to(dtype), init_out.to(dtype))
        dist.backward()
        return dist.detach(), pred.detach(), [st.param.grad.detach().clone() for st in sol.stages]
    finally:
        torch.set_default_dtype(old)


def finals(sol, data, init_out, params, logits, dtype):
    """Eval-mode chain outputs from the final parameters; `logits` (fp32 model output) is given to both
    precisions so that predict_backward is compared on identical inputs."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        for st, p in zip(sol.stages, params):
            st.param = p.to(dtype).clone()
            st.eval()
        with torch.no_grad():
            out = dict(adv=sol.forward(data.to(dtype)), pf=sol.predict_forward(init_out.to(dtype)))
            if logits is not None:
                out["pb"] = sol.predict_backward(logits.to(dtype))
            return out
    finally:
        torch.set_default_dtype(old)


def run(name):
    meta, z, data, delta0, _ = load_big(name)
    case = meta["case"]
    torch.set_num_threads(os.cpu_count() or 8)
    model = make_model(case, z)
    with torch.no_grad():
        init_out = model(data)
    replay = case["n_iter"] if case["d"] == 2 else 1
    floors = {}
    sols = {}
    for dt in (torch.float32, torch.float64):
        old = torch.get_default_dtype()
        torch.set_default_dtype(dt)           # the oracle builds its kernels / geometry in the default dtype
        sols[dt] = oracle_solver(case)
        for st in sols[dt].stages:
            st.init()
        torch.set_default_dtype(old)
    params = start_params(case, z, delta0)
    run_sol = oracle_solver(case)          # fp32 free run: the trajectory whose steps are replayed
    for st, p in zip(run_sol.stages, params):
        st.init()
        st.param = p.clone()
    for s in range(case["n_iter"]):
        if s < replay:
            d32, p32, g32 = step(sols[torch.float32], model, data, init_out, params, torch.float32)
            d64, p64, g64 = step(sols[torch.float64], model, data, init_out, params, torch.float64)
            model.to(torch.float32)
            floors["dist"] = max(floors.get("dist", 0.0), abs(float(d32) - float(d64)) / abs(float(d64)))
            floors["pred"] = max(floors.get("pred", 0.0), rel_err(p32, p64))
            for i in range(len(g32)):
                k = "grad_%d" % i
                floors[k] = max(floors.get(k, 0.0), rel_err(g32[i], g64[i]))
            print(name, "step", s, {k: "%.2e" % v for k, v in floors.items()}, flush=True)
        # advance the fp32 trajectory by one free-running step
        for st in run_sol.stages:
            st.train()
        model.zero_grad()
        dist, _, _ = run_sol.step_loss(model, data, init_out)
        dist.backward()
        for st in run_sol.stages:
            st.update(meta["steps"][0])
        params = [st.param.detach().clone() for st in run_sol.stages]
    for st in run_sol.stages:
        st.rescale()
        st.eval()
    fparams = [st.param.detach().clone() for st in run_sol.stages]
    f32 = finals(sols[torch.float32], data, init_out, fparams, None, torch.float32)
    with torch.no_grad():
        logits = model(f32["adv"])
    f32 = finals(sols[torch.float32], data, init_out, fparams, logits, torch.float32)
    f64 = finals(sols[torch.float64], data, init_out, fparams, logits, torch.float64)
    for k in f32:
        floors[k] = rel_err(f32[k], f64[k])
    print(name, "finals", {k: "%.2e" % floors[k] for k in f32}, flush=True)
    with open(os.path.join(HERE, name + ".floors.json"), "w") as f:
        json.dump(floors, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(BIG_CASES)):
        run(n)
