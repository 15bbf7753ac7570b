"""Golden fixtures at BASELINE.json's full sizes, from the UNMODIFIED reference on CPU (build container
only: needs /root/reference).

    python tests/golden/make_golden_big.py [case ...]        # default: all of cases.BIG_CASES

Same protocol as make_golden.py (the reference's own `optimizing_transform`, wrapped -- not modified -- to
record per-step start parameters, raw `param.grad` and `dist`), with two differences that keep a 128^3
case small: the volume and the AdvNoise start parameter are REGENERATED from seeds by tests/golden/bigfix.py
(the fixture records their fp64 checksums), and full-size outputs are stored as strided samples + fp64
checksums of the whole tensor.  Small tensors (bias control points, velocities, affine parameters, their
gradients, model weights) are stored in full.
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402
from tests.golden import bigfix  # noqa: E402
from tests.golden.cases import BIG_CASES  # noqa: E402
from tests.golden.make_golden import STEP, build_reference, make_model  # noqa: E402

SMALL = 1 << 16      # tensors up to this many elements are stored in full


def put(rec, key, t, case):
    t = t.detach()
    if t.numel() <= SMALL:
        rec[key] = t.clone()
    else:
        rec[key + "__s"] = bigfix.strided(t, case)
        rec[key + "__ck"] = bigfix.checksum(t)


def run_case(aug, name, case):
    t0 = time.time()
    torch.set_num_threads(os.cpu_count() or 8)
    data = bigfix.regen_data(case)
    torch.manual_seed(case["seed"])
    model = make_model(case)
    transforms = build_reference(aug, case)
    solver = aug.ComposeAdversarialTransformSolver(
        chain_of_transforms=transforms, divergence_types=["mse", "contour"],
        divergence_weights=[1.0, 0.5], use_gpu=False, if_norm_image=True)
    init_output = solver.get_init_output(model=model, data=data)
    solver.init_random_transformation()
    delta0 = bigfix.regen_delta(case)
    for t, n in zip(transforms, case["chain"]):
        if n == "noise":
            t.set_parameters(delta0)                       # injected: regenerable without the reference

    rec = {"model_w": model.weight.detach(), "model_b": model.bias.detach(),
           "data__ck": bigfix.checksum(data), "delta0__ck": bigfix.checksum(delta0)}
    put(rec, "init_output", init_output, case)
    for i, (t, n) in enumerate(zip(transforms, case["chain"])):
        if n != "noise":
            rec["p0_%d" % i] = t.param.detach().clone()

    n_iter = case["n_iter"]
    counters = {"step": [0] * len(transforms), "loss": 0}

    def wrap_update(i, t, tname):
        orig = t.optimize_parameters

        def patched(step_size=None):
            s = counters["step"][i]
            if s > 0:
                put(rec, "s%d_param_%d" % (s, i), t.param, case)
            put(rec, "s%d_grad_%d" % (s, i), t.param.grad, case)
            counters["step"][i] += 1
            return orig(step_size=step_size)
        t.optimize_parameters = patched

    for i, (t, n) in enumerate(zip(transforms, case["chain"])):
        wrap_update(i, t, n)

    orig_loss = solver.loss_fn

    def patched_loss(pred, reference, mask=None):
        val = orig_loss(pred=pred, reference=reference, mask=mask)
        s = counters["loss"]
        rec["s%d_dist" % s] = val.detach().clone()
        if s in (0, n_iter - 1):
            put(rec, "s%d_pred" % s, pred, case)
            if mask is not None:
                put(rec, "s%d_mask" % s, mask[:, :1], case)
        counters["loss"] += 1
        print("   step %d dist %.8g  (%.0f s)" % (s, float(val.detach()), time.time() - t0), flush=True)
        return val
    solver.loss_fn = patched_loss

    steps = [STEP[n] for n in case["chain"]]
    solver.optimizing_transform(model=model, data=data, init_output=init_output,
                                optimize_flags=[True] * len(transforms), n_iter=n_iter, step_sizes=steps)
    solver.loss_fn = orig_loss
    for i, t in enumerate(transforms):
        put(rec, "final_param_%d" % i, t.param, case)

    with torch.no_grad():
        adv = solver.forward(data.clone())
        logits = model(adv)
        put(rec, "adv", adv, case)
        put(rec, "logits", logits, case)
        put(rec, "pf", solver.predict_forward(init_output.clone()), case)
        put(rec, "pb", solver.predict_backward(logits.clone()), case)
    dist, _, _, _ = solver.calc_adv_consistency_loss(data.detach().clone(), model, init_output)
    rec["final_loss"] = dist.detach()
    meta = dict(case=case, steps=steps, torch=torch.__version__, reference_commit="ed5cd70", big=True,
                seconds=round(time.time() - t0, 1), threads=torch.get_num_threads())
    arrays = {}
    for k, v in rec.items():
        v = v.numpy()
        arrays[k] = np.ascontiguousarray(v if k.endswith("__ck") else v.astype(np.float32))
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-12s %7.1f KB  %.0f s  dists=%s" % (name, os.path.getsize(path) / 1024, time.time() - t0,
                                               [float(rec["s%d_dist" % s]) for s in range(n_iter)]))


def main():
    aug = ref_shim.load()
    for n in (sys.argv[1:] or list(BIG_CASES)):
        run_case(aug, n, BIG_CASES[n])


if __name__ == "__main__":
    main()
