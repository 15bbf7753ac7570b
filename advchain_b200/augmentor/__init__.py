"""Drop-in replacements for `advchain.augmentor` (augmentor/__init__.py:1-7 of the reference)."""
from .base import AdvTransformBase
from .noise import AdvNoise
from .bias import AdvBias
from .morph import AdvMorph, get_base_grid
from .affine import AdvAffine
from .solver import ComposeAdversarialTransformSolver
from .sharding import ShardContext, shard_slice

__all__ = ["AdvTransformBase", "AdvNoise", "AdvBias", "AdvMorph", "AdvAffine",
           "ComposeAdversarialTransformSolver", "get_base_grid", "ShardContext", "shard_slice"]
