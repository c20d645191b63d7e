"""thunderbolt.jl_b200 -- the B200-native monodomain hot path behind Thunderbolt.jl's solver API.

The directory name carries a dot, so import it through the repo-root shim `thunderbolt_jl_b200`.
Layout: csrc/ (sm_100a kernels + the C ABI of include/tbolt_b200.h), _lib.py (ctypes binding),
core.py (handle objects), api.py (mirror of the reference's host API), dist.py (one process per GPU).
"""
from . import _lib
from ._lib import TBError, build, declared_symbols
from .core import *  # noqa: F401,F403
from .api import *  # noqa: F401,F403
from . import core, api, multidomain, ecg, io
from .multidomain import (InterfaceDiffusionModel, PointBlockedLayout, PointwiseMultiODEFunction, StateBlock,  # noqa: F401
                          StateBlockedLayout, SubdomainGrid, insert_interfaces, state_range)
