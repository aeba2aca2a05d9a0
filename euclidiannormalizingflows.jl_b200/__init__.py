"""enf_b200 -- B200-native batched trafo-chain path of EuclidianNormalizingFlows.jl.

Host-side mirror of the reference API over libenf_b200.so (hand-written sm_100a
kernels behind a C ABI, include/enf_b200.h).  See DESIGN.md / INTEGRATION.md.
"""
from ._lib import EnfError, LIB_PATH, lib
from .device import B200Matrix, Context, default_context
from .trafos import (CenterContract, CenterStretch, ComposedFunction, HouseholderTrafo, JohnsonTrafo,
                     JohnsonTrafoInv, ScaleShiftTrafo, Trafo, compose, flatten, get_chain, inverse,
                     mvnormal_negll_trafo, mvnormal_negll_trafograd, pack_params, result_dtype, unpack_grads,
                     with_logabsdet_jacobian)
from .whitening import ADAGrad, batch_ranges, optimize_whitening, setup, update
from .johnsonsu import JohnsonSU
from .variational import GaussMixture, nELBO, nELBO_trafograd, optimise_ELBO
from . import dist

__all__ = [
    "ADAGrad", "B200Matrix", "GaussMixture", "JohnsonSU", "nELBO", "nELBO_trafograd", "optimise_ELBO", "CenterContract", "CenterStretch", "ComposedFunction", "Context", "EnfError",
    "HouseholderTrafo", "JohnsonTrafo", "JohnsonTrafoInv", "LIB_PATH", "ScaleShiftTrafo", "Trafo",
    "batch_ranges", "compose", "default_context", "dist", "flatten", "get_chain", "inverse", "lib",
    "mvnormal_negll_trafo", "mvnormal_negll_trafograd", "optimize_whitening", "pack_params", "result_dtype",
    "setup", "unpack_grads", "update", "with_logabsdet_jacobian",
]
