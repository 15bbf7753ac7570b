"""AdvAffine -- adversarial affine transformation. Drop-in for
advchain.augmentor.adv_affine.AdvAffine (adv_affine.py:12-333).

Parameters -> matrix (+ closed-form inverse) in advk_affine_theta_*; the warp generates the
affine grid on the fly inside advk_warp_affine_* (the reference materialises an N x S x d grid
with a bmm, adv_affine.py:297-298, and inverts with a batched LU, :320).
"""
import logging
import weakref

import torch

from .. import _lib
from . import _ops
from .base import AdvTransformBase

logger = logging.getLogger(__name__)


class AdvAffine(AdvTransformBase):
    def __init__(self, spatial_dims=2,
                 config_dict={'rot': 30.0 / 180.0, 'scale_x': 0.2, 'scale_y': 0.2, 'shift_x': 0.1,
                              'shift_y': 0.1, 'data_size': [1, 1, 8, 8],
                              'forward_interp': 'bilinear', 'backward_interp': 'bilinear'},
                 image_padding_mode="zeros", power_iteration=False, use_gpu=True, debug=False,
                 device=torch.device("cuda")):
        super(AdvAffine, self).__init__(spatial_dims=spatial_dims, config_dict=config_dict,
                                        use_gpu=use_gpu, debug=debug, device=device)
        self.power_iteration = power_iteration
        self.image_padding_mode = image_padding_mode
        # quirk Q3: configured interpolators only take effect after init_parameters()
        self.forward_interp = 'bilinear'
        self.backward_interp = 'bilinear'
        self.affine_matrix = None
        self._theta_entry = None
        self._inv_of = None

    def init_config(self, config_dict):
        self.translation_x = config_dict['shift_x']
        self.translation_y = config_dict['shift_y']
        self.scale_x = config_dict['scale_x']
        self.scale_y = config_dict['scale_y']
        if self.spatial_dims == 2:
            self.rot_ratio = config_dict['rot']
        if self.spatial_dims == 3:
            self.rot_x = config_dict['rot_x']
            self.rot_y = config_dict['rot_y']
            self.rot_z = config_dict['rot_z']
            self.scale_z = config_dict['scale_z']
            self.translation_z = config_dict['shift_z']
        self.xi = 1e-6
        self.data_size = config_dict['data_size']
        if 'forward_interp' in config_dict:
            self.forward_interp = config_dict['forward_interp']
        if 'backward_interp' in config_dict:
            self.backward_interp = config_dict['backward_interp']

    def _cfg(self):
        c = _lib.AffineCfg()
        c.d = self.spatial_dims
        if self.spatial_dims == 2:
            rot = [self.rot_ratio, 0.0, 0.0]
            scale = [self.scale_x, self.scale_y, 0.0]
            shift = [self.translation_x, self.translation_y, 0.0]
        else:
            rot = [self.rot_x, self.rot_y, self.rot_z]
            scale = [self.scale_x, self.scale_y, self.scale_z]
            shift = [self.translation_x, self.translation_y, self.translation_z]
        for i in range(3):
            c.rot[i], c.scale[i], c.shift[i] = float(rot[i]), float(scale[i]), float(shift[i])
        return c

    def init_parameters(self):
        self.init_config(self.config_dict)
        self.batch_size = self.data_size[0]
        self.param = self.draw_random_affine_tensor_list(batch_size=self.batch_size)
        return self.param

    def draw_random_affine_tensor_list(self, batch_size, identity_init=False):
        """adv_affine.py:166-180."""
        num_params = 5 if self.spatial_dims == 2 else 9
        if identity_init:
            return torch.zeros(batch_size, num_params, device=self.device, dtype=torch.float32)
        t = 2 * torch.rand(batch_size, num_params, dtype=torch.float32, device=self.device) - 1
        return torch.clamp(t, -1.0, 1.0)

    def train(self):
        self.is_training = True
        if self.param is None:
            self.init_parameters()
        p = self.param.detach()
        if self.power_iteration:
            p = p.sign()
        self.param = self._as_leaf(p)

    def optimize_parameters(self, step_size=None):
        """param += step * sign(grad) (adv_affine.py:182-198)."""
        if step_size is None:
            step_size = self.step_size
        p = self.param.detach().clone().float().contiguous()
        mode = _lib.UPD_SIGN_POWER if self.power_iteration else _lib.UPD_SIGN_ASCENT
        self.param = _ops.pgd_update_(p, self.param.grad, step_size, mode, guard=self._guard)
        return self.param

    def rescale_parameters(self):
        return self.param

    # ------------------------------------------------------------------ matrices
    def _pscale(self):
        return self.xi if (self.power_iteration and self.is_training) else 1.0

    def gen_batch_affine_matrix(self, affine_tensors):
        """N x (5|9) -> N x d x (d+1) (adv_affine.py:210-273)."""
        theta, _ = _ops.AffineTheta.apply(affine_tensors, self._cfg(), 1.0)
        return theta

    def _theta_pair(self):
        p = self.param
        e = self._theta_entry
        need_graph = torch.is_grad_enabled() and p.requires_grad
        if (e is not None and e[0]() is p and e[1] == p._version and e[2] == self._pscale()
                and (e[3] or not need_graph)):
            return e[4], e[5]
        theta, theta_inv = _ops.AffineTheta.apply(p, self._cfg(), self._pscale())
        self._theta_entry = (weakref.ref(p), p._version, self._pscale(), need_graph, theta, theta_inv)
        return theta, theta_inv

    def get_inverse_matrix(self, affine_matrix):
        """First d rows of inverse([theta; 0 .. 0 1]) (adv_affine.py:316-324). For the matrix produced
        by forward() the closed-form inverse from advk_affine_theta_fwd is returned; an arbitrary
        user matrix (off the hot path) is inverted with torch.linalg."""
        if self._inv_of is not None and self._inv_of[0] is affine_matrix:
            return self._inv_of[1]
        n, d = affine_matrix.shape[0], self.spatial_dims
        homo = torch.eye(d + 1, device=affine_matrix.device).repeat(n, 1, 1)
        homo[:, :d] = affine_matrix
        return torch.linalg.inv(homo)[:, :d, :]

    # ------------------------------------------------------------------ warps
    def transform(self, data, affine_matrix, interp=None, padding_mode=None):
        # quirk Q5 (adv_affine.py:293-294): a caller-supplied padding mode is always replaced by
        # the transform's own image_padding_mode
        padding_mode = self.image_padding_mode
        if interp is None:
            interp = self.forward_interp
        pad, padv = _ops.parse_padding(padding_mode, data)
        return _ops.WarpAffine.apply(data, affine_matrix, pad, _ops.parse_interp(interp), padv)

    def _stage(self, mode, data, interp=None, padding_mode=None):
        if self.param is None:
            self.init_parameters()
        fwd = mode in ("fwd", "pfwd")
        if interp is None:
            interp = self.forward_interp if fwd else self.backward_interp
        if _ops.parse_interp(interp) == _ops._lib.INTERP_BICUBIC:
            return NotImplemented                                     # bicubic: per-transform kernels
        pp = _ops.fused_padding(self.image_padding_mode, data)      # quirk Q5
        if pp is None:
            return NotImplemented
        theta, theta_inv = self._theta_pair()
        if fwd:
            self.affine_matrix = theta
            self._inv_of = (theta, theta_inv)
        else:
            assert self.affine_matrix is not None, 'play forward before backward'
            if self._inv_of is None or self._inv_of[0] is not self.affine_matrix:
                return NotImplemented                                 # user-replaced matrix: generic path
            theta_inv = self._inv_of[1]
        return dict(kind="affine", theta=theta if fwd else theta_inv, pad=pp[0], padv=pp[1],
                    interp=_ops.parse_interp(interp))

    def forward(self, data, interp=None, padding_mode=None):
        """adv_affine.py:121-146."""
        if self.param is None:
            self.init_parameters()
        if interp is None:
            interp = self.forward_interp
        theta, theta_inv = self._theta_pair()
        self.affine_matrix = theta
        self._inv_of = (theta, theta_inv)
        out = self.transform(data, theta, interp=interp, padding_mode=padding_mode)
        src = data
        self.diff = lambda: src - out.detach()
        return out

    def backward(self, data, interp=None, padding_mode=None):
        """adv_affine.py:154-164."""
        assert self.param is not None and self.affine_matrix is not None, 'play forward before backward'
        if interp is None:
            interp = self.backward_interp
        inverse_matrix = self.get_inverse_matrix(self.affine_matrix)
        return self.transform(data, inverse_matrix, interp=interp, padding_mode=padding_mode)

    def predict_forward(self, data, interp=None, padding_mode=None):
        return self.forward(data, interp=interp, padding_mode=padding_mode)

    def predict_backward(self, data, interp=None, padding_mode=None):
        return self.backward(data, interp=interp, padding_mode=padding_mode)

    def get_name(self):
        return 'affine'

    def is_geometric(self):
        return 1
