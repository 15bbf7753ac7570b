"""AdvNoise -- additive adversarial noise. Drop-in for advchain.augmentor.adv_noise.AdvNoise
(adv_noise.py:10-119); device work in advk_intensity_* / advk_pgd_update."""
import logging

import torch

from . import _ops
from .base import AdvTransformBase

logger = logging.getLogger(__name__)


class AdvNoise(AdvTransformBase):
    def __init__(self, spatial_dims=2,
                 config_dict={'epsilon': 0.1, 'xi': 1e-6, 'data_size': [10, 1, 8, 8]},
                 power_iteration=False, ignore_values=None, use_gpu=True, debug=False,
                 device=torch.device("cuda")):
        super(AdvNoise, self).__init__(spatial_dims=spatial_dims, config_dict=config_dict,
                                       use_gpu=use_gpu, debug=debug, device=device)
        self.power_iteration = power_iteration
        self.ignore_values = ignore_values

    def init_config(self, config_dict):
        self.epsilon = config_dict['epsilon']
        self.xi = config_dict['xi']
        self.data_size = config_dict['data_size']

    def init_parameters(self):
        """unit-L2 Gaussian noise per sample (adv_noise.py:41-49)."""
        noise = self.unit_normalize(torch.randn(*self.data_size, device=self.device, dtype=torch.float32))
        self.param = noise
        return noise

    def train(self):
        self.is_training = True
        if self.param is None:
            self.init_parameters()
        p = self.param.detach()
        if self.power_iteration:
            p = self.unit_normalize(p)
        self.param = self._as_leaf(p)

    def _scale(self):
        return self.xi if (self.power_iteration and self.is_training) else self.epsilon

    def forward(self, data, **kwargs):
        """x + eps*delta (xi*delta in power-iteration training), adv_noise.py:67-90."""
        if self.param is None:
            self.init_parameters()
        out = _ops.Intensity.apply(data, self.param, None, _ops.ORDER_NOISE, self._scale(), None, 1.0,
                                   self.ignore_values)
        src = data
        self.diff = lambda: out.detach() - src
        return out

    def _stage(self, mode, data, interp=None, padding_mode=None):
        """Stage of the fused chain (None: identity for the prediction paths)."""
        if mode != "fwd":
            return None
        if self.param is None:
            self.init_parameters()
        return dict(kind="intensity", order=_ops.ORDER_NOISE, noise_scale=self._scale(),
                    ignore=self.ignore_values, delta=self.param)

    def optimize_parameters(self, step_size=None):
        if step_size is None:
            step_size = self.step_size
        return self._l2_step(step_size)

    def rescale_parameters(self):
        self.param = self.unit_normalize(self.param, p_type='l2')

    def backward(self, data, **kwargs):
        return data

    def predict_forward(self, data, **kwargs):
        return data

    def predict_backward(self, data, **kwargs):
        return data

    def get_name(self):
        return 'noise'
