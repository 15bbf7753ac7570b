"""Import shim for the UNMODIFIED reference (build container only).  TEST INFRASTRUCTURE.

`/root/reference` exists only in the build container, never on the GPU box, so this module is
used exclusively by `tests/golden/make_golden.py` (fixture generation) and by optional CPU
tests that skip when the reference is absent.  Nothing is copied: the reference package is
imported from where it lies.

Blockers worked around: `matplotlib`, `SimpleITK` are not installed but
imported at module top level (adv_affine.py:2, common/utils.py:7, common/vis.py:1); `np.Inf`
was removed in NumPy 2 (adv_bias.py:237-238).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ADVCHAIN_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "advchain", "augmentor")):
    # GPU box: the unmodified package mirrored by __graft_entry__.install_reference() (bench.py's reference arms)
    _alt = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if os.path.isdir(os.path.join(_alt, "advchain", "augmentor")):
        REFERENCE_ROOT = _alt


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "advchain", "augmentor"))


def load():
    """Returns the reference's `advchain.augmentor` module."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    for name in ("matplotlib", "matplotlib.pyplot", "SimpleITK"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import advchain.augmentor as aug
    return aug
