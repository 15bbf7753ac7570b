"""Golden-vector case definitions shared by `make_golden.py` (which runs the unmodified
reference in the build container) and the tests that replay the fixtures.

Sizes are deliberately non-square / non-cubic so that an x/y/z axis mix-up cannot pass.
Hyper-parameters follow the reference README / notebooks (README.md:101-143) scaled to the
fixture sizes: bias spacing = size//2, morph vector = size//16 (rounded up to >=2).
"""


def _affine_cfg(d, size):
    if d == 2:
        return {"rot": 30.0 / 180.0, "scale_x": 0.2, "scale_y": 0.2, "shift_x": 0.1, "shift_y": 0.1,
                "data_size": size, "forward_interp": "bilinear", "backward_interp": "bilinear"}
    return {"rot_x": 10.0 / 180.0, "rot_y": 10.0 / 180.0, "rot_z": 10.0 / 180.0,
            "scale_x": 0.1, "scale_y": 0.1, "scale_z": 0.1,
            "shift_x": 0.1, "shift_y": 0.1, "shift_z": 0.1,
            "data_size": size, "forward_interp": "bilinear", "backward_interp": "bilinear"}


def stage_cfgs(d, size, vector=None, spacing=None, downscale=None):
    sp = list(size[2:])
    return {
        "noise": {"epsilon": 1.0, "xi": 1e-6, "data_size": size},
        "bias": {"epsilon": 0.3, "control_point_spacing": spacing or [s // 2 for s in sp],
                 "downscale": downscale or (2 if d == 2 else 4), "data_size": size,
                 "interpolation_order": 3, "init_mode": "random", "space": "log"},
        "morph": {"epsilon": 1.5, "data_size": size,
                  "vector_size": vector or [max(2, s // 16) for s in sp]},
        "affine": _affine_cfg(d, size),
    }


FULL = ["noise", "bias", "morph", "affine"]

CASES = {
    # name: dict(d, size, chain, n_iter, K, padding, seed)
    "c2d_full": dict(d=2, size=[2, 1, 48, 64], chain=FULL, n_iter=3, K=4, seed=11),
    "c2d_affine": dict(d=2, size=[1, 1, 40, 56], chain=["affine"], n_iter=2, K=4, seed=12),
    "c2d_morph_border": dict(d=2, size=[2, 1, 32, 48], chain=["morph", "affine"], n_iter=2, K=3,
                             seed=13, padding="border"),
    "c2d_numeric_pad": dict(d=2, size=[2, 2, 32, 32], chain=["bias", "noise", "affine", "morph"],
                            n_iter=2, K=4, seed=14, padding=-1.0),
    "c2d_nogeo": dict(d=2, size=[2, 1, 32, 48], chain=["noise", "bias"], n_iter=2, K=4, seed=15),
    "c3d_full": dict(d=3, size=[2, 1, 16, 24, 32], chain=FULL, n_iter=2, K=4, seed=21,
                     vector=[2, 3, 4]),
    "c3d_morph_affine": dict(d=3, size=[1, 1, 24, 16, 32], chain=["morph", "affine"], n_iter=2,
                             K=4, seed=22, vector=[3, 2, 4]),
    "c3d_affine": dict(d=3, size=[2, 1, 12, 20, 16], chain=["affine"], n_iter=1, K=2, seed=23),
    # BASELINE.json configs[0]: 2-D 1x1x192x192, AdvAffine only, 1 PGD step (the notebook path)
    "c1_affine192": dict(d=2, size=[1, 1, 192, 192], chain=["affine"], n_iter=1, K=4, seed=51),
    # power-iteration (VAT-style) training mode of every transform: xi-scaled probes, g -> unit(g)
    "c2d_power": dict(d=2, size=[2, 1, 40, 48], chain=FULL, n_iter=2, K=4, seed=31, power=True),
    "c3d_power": dict(d=3, size=[1, 1, 16, 20, 24], chain=FULL, n_iter=2, K=3, seed=32, power=True,
                      vector=[2, 2, 3]),
    # anatomy-preserving branch of optimizing_transform (adv_compose_solver.py:329-338, 376-400): the
    # thresholded round-trip score enters dist and the stop test; tolerance 1.0 -> exactly n_iter steps
    "c2d_anatomy": dict(d=2, size=[2, 1, 40, 48], chain=["morph", "affine"], n_iter=1, K=4, seed=41,
                        anatomy=True),
}

# Full-size cases = BASELINE.json's configurations at their stated sizes (run by make_golden_big.py).  The
# inputs (volume, noise parameter) are regenerated from seeds and verified by fp64 checksums; outputs are
# stored as strided samples (stride coprime with every tile size of the kernels) + fp64 checksums of the
# whole tensor, so that a 128^3 case stays a few hundred KB.
BIG_CASES = {
    # BASELINE.json configs[1]: 2-D 8x1x256x256, full chain, 5 PGD steps
    "c2_full256": dict(d=2, size=[8, 1, 256, 256], chain=FULL, n_iter=5, K=4, seed=61, stride=5, offset=2),
    # BASELINE.json configs[2]: 3-D 2x1x128x128x64, full chain, 5 PGD steps
    "c3_full128": dict(d=3, size=[2, 1, 128, 128, 64], chain=FULL, n_iter=5, K=4, seed=62, stride=5, offset=2),
    # the metric line: 3-D 1x1x128^3, full chain, one PGD iteration
    "m128_step": dict(d=3, size=[1, 1, 128, 128, 128], chain=FULL, n_iter=1, K=4, seed=63, stride=5, offset=2),
}
