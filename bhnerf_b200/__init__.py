"""bhnerf_b200: B200-native (sm_100a) implementation of bhnerf's differentiable renderer + train step.

Module names mirror the reference package (`network`, `emission`, `kgeo`, `optimization`, `constants`,
`utils`) for the hot path only; everything computes through libbhnerf_b200.so (include/bhnerf_b200.h)."""
from . import constants, utils  # noqa: F401
from . import engine  # noqa: F401
from . import emission, kgeo, network, optimization  # noqa: F401

__version__ = '0.1.0'
