"""advchain_b200 -- Blackwell-native implementation of advchain's chained adversarial
augmentation hot path (AdvNoise, AdvBias, AdvMorph, AdvAffine, ComposeAdversarialTransformSolver).

    from advchain_b200.augmentor import AdvNoise, AdvBias, AdvMorph, AdvAffine, \
        ComposeAdversarialTransformSolver

or, to run unmodified user code written against the reference package:

    import advchain_b200; advchain_b200.install_as_advchain()
    from advchain.augmentor import ...          # now resolves to this package
"""
import sys

__version__ = "0.1.0"


def install_as_advchain():
    """Registers this package under the reference's import names (`advchain.augmentor`,
    `advchain.common.loss`, `advchain.common.utils`, `advchain.common.layers`)."""
    import types
    from . import augmentor, common
    from .common import layers, loss, utils
    root = types.ModuleType("advchain")
    root.augmentor, root.common = augmentor, common
    sys.modules["advchain"] = root
    sys.modules["advchain.augmentor"] = augmentor
    sys.modules["advchain.common"] = common
    sys.modules["advchain.common.loss"] = loss
    sys.modules["advchain.common.utils"] = utils
    sys.modules["advchain.common.layers"] = layers
    # the reference's per-class modules (user code imports e.g. advchain.augmentor.adv_bias.AdvBias)
    from .augmentor import affine, base, bias, morph, noise, solver
    for name, mod in (("adv_affine", affine), ("adv_bias", bias), ("adv_morph", morph), ("adv_noise", noise),
                      ("adv_transformation_base", base), ("adv_compose_solver", solver)):
        sys.modules["advchain.augmentor." + name] = mod
        setattr(augmentor, name, mod)
    return root
