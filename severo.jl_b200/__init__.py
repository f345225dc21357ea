"""severo.jl_b200 — B200-native implementation of Severo.jl's IRLBA-PCA hot path.

Layout: ``csrc/`` hand-written sm_100a CUDA kernels behind the C ABI of include/severo_b200.h,
``api.py`` the host-side mirror of the reference's Julia interface, ``sharding.py`` the cell sharding
for N GPUs. Import as ``severo_jl_b200`` (shim at the repo root; the directory name has a dot).
"""
from ._lib import LIB_PATH, SIGNATURES, SeveroB200Error, init, lib, load  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import _pca, variance_stabilizing_transformation  # noqa: F401
from . import sharding  # noqa: F401
