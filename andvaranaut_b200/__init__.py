"""andvaranaut_b200 -- B200-native Gaussian-process inner loop behind andvaranaut's GPMCMC surrogate API.

``from andvaranaut_b200 import *`` gives the names the reference's ``from andvaranaut import *`` gives for this
path: ``GPMCMC``, ``LHC``, the transform classes, ``save_object`` / ``load_object``.
"""
from .core import save_object, load_object  # noqa: F401
from .lhc import LHC  # noqa: F401
from .transform import *  # noqa: F401,F403
from .gpmcmc import GPMCMC  # noqa: F401
