"""AdvTransformBase -- the contract every adversarial transform implements.

Drop-in for `advchain.augmentor.adv_transformation_base.AdvTransformBase`
(adv_transformation_base.py:5-189): same constructor, same method names, same public
attributes (`param, is_training, power_iteration, diff, step_size, data_size, device`), so
user subclasses (README.md:293-294) keep working inside the solver.
"""
import torch

from .. import _lib
from . import _ops


class AdvTransformBase(object):
    def __init__(self, spatial_dims=2, config_dict={"data_size": [1, 1, 1, 1]}, use_gpu=True,
                 device=torch.device("cuda"), debug=False):
        self.spatial_dims = spatial_dims
        assert self.spatial_dims == 2 or self.spatial_dims == 3, 'only support 2D/3D'
        self.config_dict = config_dict
        data_dim = len(config_dict["data_size"])
        assert data_dim == self.spatial_dims + 2, \
            f"check data size in the config file, should be {self.spatial_dims+2}D, but got {data_dim}D"
        self.param = None
        self.is_training = False
        self.use_gpu = use_gpu
        self.device = torch.device(device) if self.use_gpu else torch.device("cpu")
        self.debug = debug
        self._diff = None
        self._guard = None          # device scalar: skip the update when it is NaN/Inf (graph mode)
        self.power_iteration = False
        self.init_config(self.config_dict)
        self.step_size = 1

    # -- the reference materialises `diff` in every forward (an extra full-size pass, quirk Q11);
    # here it is a lazily evaluated attribute so the hot loop never pays for it.
    @property
    def diff(self):
        d = self._diff
        if callable(d):
            d = d()
            self._diff = d
        return d

    @diff.setter
    def diff(self, value):
        self._diff = value

    def init_config(self, config_dict):
        raise NotImplementedError

    def init_parameters(self):
        raise NotImplementedError

    def set_parameters(self, param):
        self.param = param.detach().clone()

    def get_parameters(self):
        return self.param

    def set_step_size(self, step_size=1):
        self.step_size = step_size

    def get_step_size(self):
        return self.step_size

    def _as_leaf(self, p):
        """Fresh autograd leaf (the reference re-wraps the parameter as nn.Parameter every step)."""
        return torch.nn.Parameter(p.detach(), requires_grad=True)

    def train(self):
        if self.param is None:
            self.init_parameters()
        self.is_training = True
        self.param = self._as_leaf(self.param.clone())

    def eval(self):
        if self.is_training:
            self.param = self.param.detach()
            self.is_training = False

    def rescale_parameters(self, param=None):
        if param is None:
            param = self.param
        self.param = param.renorm(p=2, dim=0, maxnorm=self.epsilon)
        return self.param

    def optimize_parameters(self, step_size=None):
        raise NotImplementedError

    def forward(self, data, **kwargs):
        raise NotImplementedError

    def backward(self, data, **kwargs):
        raise NotImplementedError

    def predict_forward(self, data, **kwargs):
        raise NotImplementedError

    def predict_backward(self, data, **kwargs):
        raise NotImplementedError

    def unit_normalize(self, d, p_type='l2'):
        """Per-sample normalisation (adv_transformation_base.py:129-156). The L2 case -- the only
        one on the PGD path -- runs through advk_pgd_update; l1 / infinity are init-time helpers."""
        if p_type == 'l2':
            if d.is_cuda:
                out = d.detach().clone().float().contiguous()
                return _ops.pgd_update_(out, out, 0.0, _lib.UPD_L2_POWER)
            flat = d.reshape(d.shape[0], -1)
            nrm = flat.norm(dim=1).view(-1, *([1] * (d.dim() - 1)))
            return d / (nrm + 1e-20)
        flat = d.reshape(d.size(0), -1)
        if p_type == 'l1':
            return (flat / flat.norm(p=1, dim=1, keepdim=True)).view(d.size())
        if p_type == 'infinity':
            return (flat / (1e-20 + torch.max(flat, 1, keepdim=True)[0])).view(d.size())
        return d

    def rescale_intensity(self, data, new_min=0, new_max=1, eps=1e-20):
        flat = data.reshape(data.size(0) * data.size(1), -1)
        hi = flat.max(dim=1, keepdim=True).values
        lo = flat.min(dim=1, keepdim=True).values
        return ((flat - lo + eps) / (hi - lo + eps) * (new_max - new_min) + new_min).view(data.size())

    def get_name(self):
        raise NotImplementedError

    def is_geometric(self):
        return 0

    # helpers shared by the concrete transforms -------------------------------------------
    def _l2_step(self, step_size):
        """param <- param + step * g/||g||  (or g/||g|| in power iteration), on device, in one call."""
        grad = self.param.grad
        p = self.param.detach().clone().float().contiguous()
        mode = _lib.UPD_L2_POWER if self.power_iteration else _lib.UPD_L2_ASCENT
        self.param = _ops.pgd_update_(p, grad, 0.0 if self.power_iteration else step_size, mode,
                                      guard=self._guard)
        return self.param
