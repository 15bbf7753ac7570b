"""Model-side context managers the solver uses (reference: advchain/common/utils.py:115-173).

They only toggle attributes on the user's model; nothing here touches the device path.
"""
import contextlib

import torch


def _is_fixable_dropout(module):
    # the reference tests isinstance(Fixable2DDropout/Fixable3DDropout) (common/layers.py);
    # duck-typed here so that models built with the reference's own layer classes still match
    return type(module).__name__.startswith("Fixable") and hasattr(module, "lazy_load")


@contextlib.contextmanager
def _disable_tracking_bn_stats(model):
    """Freeze BatchNorm running-stat updates (and flip fixable dropout) inside the block."""
    saved = {}
    for name, m in model.named_modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            saved[name] = m.track_running_stats
            m.track_running_stats = False
        if _is_fixable_dropout(m):
            m.lazy_load = not m.lazy_load
    try:
        yield
    finally:
        for name, m in model.named_modules():
            if name in saved:
                m.track_running_stats = saved[name]
            if _is_fixable_dropout(m):
                m.lazy_load = not m.lazy_load


@contextlib.contextmanager
def _fix_dropout(model):
    """Replay the last dropout mask of every fixable dropout layer inside the block."""
    for _, m in model.named_modules():
        if _is_fixable_dropout(m):
            m.lazy_load = not m.lazy_load
    try:
        yield
    finally:
        for _, m in model.named_modules():
            if _is_fixable_dropout(m):
                m.lazy_load = not m.lazy_load


def set_grad(module, requires_grad=False):
    for p in module.parameters():
        p.requires_grad = requires_grad
