"""Generate the golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py c2d_full   # one case

For every case in `cases.py` it builds the reference transforms + solver (fp32, CPU, seeded),
lets the reference draw its own random parameters, runs the reference's own
`optimizing_transform` PGD loop, and records -- through wrappers around the reference objects,
the reference code itself is not modified -- for every inner step the parameters the step
started from, the raw `param.grad` the reference's autograd produced, and the scalar `dist`;
then the final parameters, `solver.forward(data)`, `predict_forward`, `predict_backward`,
the valid-region mask and the final consistency loss.  Replaying a step from the recorded
start parameters ("teacher forcing") is what the parity tests do, because free-running PGD
amplifies 1e-6 differences.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402
from tests.golden.cases import CASES, stage_cfgs  # noqa: E402

STEP = {"noise": 1.0, "bias": 1.0, "morph": 1.0, "affine": 0.1}


def build_reference(aug, case):
    d, size = case["d"], case["size"]
    cfgs = stage_cfgs(d, size, vector=case.get("vector"))
    pad = case.get("padding", "zeros")
    cpu = torch.device("cpu")
    out = []
    for name in case["chain"]:
        kw = dict(spatial_dims=d, config_dict=cfgs[name], use_gpu=False, device=cpu,
                  power_iteration=bool(case.get("power", False)))
        if name == "noise":
            out.append(aug.AdvNoise(**kw))
        elif name == "bias":
            out.append(aug.AdvBias(**kw))
        elif name == "morph":
            out.append(aug.AdvMorph(image_padding_mode=pad, **kw))
        else:
            out.append(aug.AdvAffine(image_padding_mode=pad, **kw))
    return out


def make_model(case):
    conv = torch.nn.Conv2d if case["d"] == 2 else torch.nn.Conv3d
    return conv(case["size"][1], case["K"], 3, 1, 1).eval()


def run_case(aug, name, case):
    torch.manual_seed(case["seed"])
    torch.set_num_threads(8)
    data = torch.rand(*case["size"])
    model = make_model(case)
    transforms = build_reference(aug, case)
    solver = aug.ComposeAdversarialTransformSolver(
        chain_of_transforms=transforms, divergence_types=["mse", "contour"],
        divergence_weights=[1.0, 0.5], use_gpu=False, if_norm_image=True)
    init_output = solver.get_init_output(model=model, data=data)
    solver.init_random_transformation()

    rec = {"data": data, "model_w": model.weight.detach(), "model_b": model.bias.detach(),
           "init_output": init_output}
    for i, t in enumerate(transforms):
        rec["p0_%d" % i] = t.param.detach().clone()

    counters = {"step": [0] * len(transforms), "loss": 0}

    def wrap_update(i, t):
        orig = t.optimize_parameters

        def patched(step_size=None):
            s = counters["step"][i]
            rec["s%d_param_%d" % (s, i)] = t.param.detach().clone()
            rec["s%d_grad_%d" % (s, i)] = t.param.grad.detach().clone()
            counters["step"][i] += 1
            return orig(step_size=step_size)
        t.optimize_parameters = patched

    for i, t in enumerate(transforms):
        wrap_update(i, t)

    orig_loss = solver.loss_fn

    def patched_loss(pred, reference, mask=None):
        val = orig_loss(pred=pred, reference=reference, mask=mask)
        s = counters["loss"]
        rec["s%d_dist" % s] = val.detach().clone()
        if s == 0:
            rec["s0_pred"] = pred.detach().clone()
            if mask is not None:
                rec["s0_mask"] = mask.detach().clone()
        counters["loss"] += 1
        return val
    solver.loss_fn = patched_loss

    steps = [STEP[n] for n in case["chain"]]
    extra = {}
    if case.get("anatomy"):
        amask = (data > 0.5).float()
        rec["anatomy_mask"] = amask
        orig_anat = solver.compute_anatomy_misoverlapping_loss
        scores = []

        def patched_anat(anatomy_mask_images):
            v = orig_anat(anatomy_mask_images=anatomy_mask_images)
            scores.append(float(v))
            return v
        solver.compute_anatomy_misoverlapping_loss = patched_anat
        extra = dict(anatomy_mask_images=amask, anatomy_reg_weight=50, volume_preserve_tolerance=1.0)
    solver.optimizing_transform(model=model, data=data, init_output=init_output,
                                optimize_flags=[True] * len(transforms), n_iter=case["n_iter"],
                                step_sizes=steps, **extra)
    if case.get("anatomy"):
        solver.compute_anatomy_misoverlapping_loss = orig_anat
        rec["anatomy_scores"] = torch.tensor(scores)
    solver.loss_fn = orig_loss
    for i, t in enumerate(transforms):
        rec["final_param_%d" % i] = t.param.detach().clone()

    with torch.no_grad():
        adv = solver.forward(data.clone())
        logits = model(adv)
        rec["adv"] = adv
        rec["logits"] = logits
        rec["pf"] = solver.predict_forward(init_output.clone())
        rec["pb"] = solver.predict_backward(logits.clone())
    dist, _, _, _ = solver.calc_adv_consistency_loss(data.detach().clone(), model, init_output)
    rec["final_loss"] = dist.detach()
    meta = dict(case=case, steps=steps, torch=torch.__version__, reference_commit="ed5cd70")
    arrays = {k: np.ascontiguousarray(v.numpy().astype(np.float32)) for k, v in rec.items()}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-18s %6.1f KB  dists=%s" % (name, os.path.getsize(path) / 1024,
                                        [float(rec["s%d_dist" % s]) for s in range(case["n_iter"])]))


def main():
    aug = ref_shim.load()
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(aug, n, CASES[n])


if __name__ == "__main__":
    main()
