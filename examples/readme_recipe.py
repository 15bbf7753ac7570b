"""The reference README's usage recipe (README.md:101-214 of cherise215/advchain), unchanged except for
the first two lines: the four transforms, the composed solver, one adversarial_training call inside a
training step.  Needs a B200 (sm_100a) and the built library (`python -m advchain_b200.build`).

    python examples/readme_recipe.py [--dims 3] [--size 64]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import advchain_b200  # noqa: E402

advchain_b200.install_as_advchain()          # `import advchain...` now resolves to the B200 implementation

import torch  # noqa: E402
from advchain.augmentor import (AdvAffine, AdvBias, AdvMorph, AdvNoise,  # noqa: E402
                                ComposeAdversarialTransformSolver)
from advchain.common.utils import random_chain  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, default=2)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
d, n, k = a.dims, 4, 4
spatial = [a.size] * d
data_size = [n, 1] + spatial
images = torch.rand(*data_size, device=dev)
labels = torch.randint(0, k, [n] + spatial, device=dev)
conv = torch.nn.Conv2d if d == 2 else torch.nn.Conv3d
model = torch.nn.Sequential(conv(1, 16, 3, padding=1), torch.nn.ReLU(), conv(16, k, 3, padding=1)).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=1e-2)

noise = AdvNoise(spatial_dims=d, config_dict={'epsilon': 1, 'xi': 1e-6, 'data_size': data_size}, device=dev)
bias = AdvBias(spatial_dims=d, config_dict={'epsilon': 0.3, 'control_point_spacing': [s // 2 for s in spatial],
                                            'downscale': 2 if d == 2 else 4, 'data_size': data_size,
                                            'interpolation_order': 3, 'init_mode': 'random', 'space': 'log'}, device=dev)
morph = AdvMorph(spatial_dims=d, config_dict={'epsilon': 1.5, 'data_size': data_size,
                                              'vector_size': [max(2, s // 16) for s in spatial]}, device=dev)
if d == 2:
    acfg = {'rot': 30 / 180.0, 'scale_x': 0.2, 'scale_y': 0.2, 'shift_x': 0.1, 'shift_y': 0.1,
            'data_size': data_size, 'forward_interp': 'bilinear', 'backward_interp': 'bilinear'}
else:
    acfg = {'rot_x': 10 / 180.0, 'rot_y': 10 / 180.0, 'rot_z': 10 / 180.0, 'scale_x': 0.1, 'scale_y': 0.1,
            'scale_z': 0.1, 'shift_x': 0.1, 'shift_y': 0.1, 'shift_z': 0.1, 'data_size': data_size,
            'forward_interp': 'bilinear', 'backward_interp': 'bilinear'}
affine = AdvAffine(spatial_dims=d, config_dict=acfg, device=dev)

for it in range(a.iters):
    chain = random_chain([noise, bias, morph, affine], max_length=4)        # a random sub-chain per iteration
    solver = ComposeAdversarialTransformSolver(chain_of_transforms=chain, divergence_types=['mse', 'contour'],
                                               divergence_weights=[1.0, 0.5], use_gpu=True, debug=False,
                                               if_norm_image=True)
    solver.use_cuda_graph = True                                             # extra: replay the PGD step from a CUDA graph
    model.train()
    opt.zero_grad()
    sup = torch.nn.functional.cross_entropy(model(images), labels)
    reg = solver.adversarial_training(data=images, model=model, n_iter=1, lazy_load=[False] * len(chain),
                                      optimize_flags=[True] * len(chain), step_sizes=1)
    (sup + reg).backward()
    opt.step()
    print("iter %d  chain %-28s supervised %.4f  adversarial consistency %.6f"
          % (it, "->".join(t.get_name() for t in chain), sup.item(), reg.item()))
print("ok")
