"""Dropout layers whose mask can be replayed (reference: advchain/common/layers.py).

The solver needs the user's model to apply the SAME dropout mask to the clean and to the perturbed
input (`_fix_dropout` / `_disable_tracking_bn_stats` flip `lazy_load` around the second forward pass).
A layer remembers the RNG seed of its last mask; with `lazy_load` it re-seeds with it instead of
drawing a new one.  Model-side helper only -- nothing here is on the device hot path."""
import torch
from torch.nn import functional as F


class _FixableDropout(torch.nn.Module):
    _fn = None

    def __init__(self, p=0.5, inplace=False, lazy_load=False, training=True):
        super().__init__()
        if p < 0 or p > 1:
            raise ValueError("dropout probability has to be between 0 and 1, but got {}".format(p))
        self.p = p
        self.inplace = inplace
        self.seed = None
        self.lazy_load = lazy_load
        self.training = training

    def forward(self, x):
        replay = self.training and self.lazy_load and self.seed is not None
        seed = self.seed if replay else torch.seed()
        self.seed = seed
        torch.manual_seed(seed)
        return type(self)._fn(x, p=self.p, training=self.training, inplace=self.inplace)


class Fixable2DDropout(_FixableDropout):
    """Channel dropout for N x C x H x W inputs with a replayable mask."""
    _fn = staticmethod(F.dropout2d)


class Fixable3DDropout(_FixableDropout):
    """Channel dropout for N x C x D x H x W inputs with a replayable mask."""
    _fn = staticmethod(F.dropout3d)
