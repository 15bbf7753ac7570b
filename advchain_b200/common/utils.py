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


def random_chain(alist, max_length=None, size_list=None):
    """A random-length prefix of a random permutation of `alist` (reference:
    advchain/common/utils.py:180-212, the README training recipe samples a sub-chain of transforms per
    iteration with it).  `size_list`, when given, is permuted identically and cut to the same length.

    The reference implementation passes a second argument to `random.shuffle`, which Python >= 3.11
    rejects, and its single-element branch reads an undefined name; this version keeps the intended
    behaviour (uniform length in [1, max_length], one shared permutation) on every Python version.
    Like the reference it permutes the input lists in place."""
    import random

    import numpy as np
    length = len(alist)
    assert length >= 1, "input list must contains at least one element"
    max_length = length if max_length is None else min(max_length, length)
    if length == 1:
        if size_list is not None:
            assert len(size_list) == 1, "must share equal size"
            return [alist[0]], [size_list[0]]
        return [alist[0]]
    sub_len = int(np.random.randint(low=1, high=max_length + 1))
    perm = list(range(length))
    random.shuffle(perm)
    alist[:] = [alist[i] for i in perm]
    if size_list is not None:
        assert len(size_list) == length, "must share equal size"
        size_list[:] = [size_list[i] for i in perm]
        return alist[:sub_len], size_list[:sub_len]
    return alist[:sub_len]
