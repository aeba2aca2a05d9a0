"""Variational objective of examples/nf_variational_1d.jl on the fused loss+gradient kernel.

  nELBO(trafo, xi)             examples/nf_variational_1d.jl:29-41
  nELBO_trafograd(trafo, xi)   :43-47
  optimise_ELBO(...)           :49-69

`xi` is a D x N matrix of standard-normal draws with the samples as COLUMNS, like every sample matrix of this package.
The example builds `vcat(xi, -xi)` of shape (2 batchsize) x 1 and lets the length-1 parameter vectors broadcast along
its rows (:32-34 swap the roles of rows and columns); `optimise_ELBO` here makes the same antithetic pairs as a
1 x (2 batchsize) matrix.  The target log-density is a fixed family (no callbacks cross the C ABI): a Gaussian mixture
applied element-wise, of which the example's `my_ll` is the default instance.
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from .device import B200Matrix
from .trafos import _as_device, get_chain, unpack_grads
from .whitening import ADAGrad, setup, update


@dataclass
class GaussMixture:
    """log p(z) = log sum_k weights[k] N(z | means[k], sigmas[k]); default: my_ll of examples/nf_variational_1d.jl:25-27."""
    weights: Sequence[float] = (0.3, 0.5, 0.2)
    means: Sequence[float] = (2.0, 5.0, -1.0)
    sigmas: Sequence[float] = (1.0, 1.0, 1.0)

    def _c(self):
        K = len(self.weights)
        if not (len(self.means) == len(self.sigmas) == K):
            raise ValueError("weights, means and sigmas need the same length")
        arrs = [(C.c_double * K)(*[float(v) for v in a]) for a in (self.weights, self.means, self.sigmas)]
        t = L.enf_target(L.ENF_TARGET_GAUSS_MIXTURE, K, *[C.cast(a, C.POINTER(C.c_double)) for a in arrs])
        return t, arrs      # keep the arrays alive for the call


def _elbo(trafo, xi, target, want_grad: bool, zygote_primal: bool):
    X = _as_device(xi, trafo)
    ch = get_chain(trafo, X.D, X.dtype, X.ctx)
    t, keep = (target or GaussMixture())._c()
    out = C.c_double()
    g = np.empty(ch.nparams, dtype=X.dtype) if want_grad else None
    L.check(X.ctx._lib.enf_elbo_grad(ch.handle, C.byref(t), C.c_void_p(X.ptr), X.N, L.ENF_NEGLL_ZYGOTE_PRIMAL if zygote_primal else 0,
                                     C.byref(out), g.ctypes.data_as(C.c_void_p) if want_grad else None), X.ctx.handle)
    del keep
    return out.value, (unpack_grads(trafo, g, X.D) if want_grad else None)


def nELBO(trafo, xi, target: Optional[GaussMixture] = None) -> float:
    """examples/nf_variational_1d.jl:29-41 for xi of shape D x N (samples are columns)."""
    return _elbo(trafo, xi, target, False, False)[0]


def nELBO_trafograd(trafo, xi, target: Optional[GaussMixture] = None, *, zygote_primal: bool = True):
    """examples/nf_variational_1d.jl:43-47 -> (nelbo, d_trafo).  zygote_primal as for mvnormal_negll_trafograd: under
    Zygote the ScaleShiftTrafo ladj VALUE is zero (src/abstract_trafo.jl:30-33), the gradient is unaffected."""
    return _elbo(trafo, xi, target, True, zygote_primal)


def optimise_ELBO(initial_trafo, optimizer: ADAGrad, *, target: Optional[GaussMixture] = None, batchsize: int = 100,
                  nepochs: int = 100, optstate=None, nelbo_history=None, rng: Optional[np.random.Generator] = None,
                  batches=None, dtype=np.float64, ctx=None):
    """examples/nf_variational_1d.jl:49-69: per epoch `batchsize` standard-normal draws, antithetic pairs
    `vcat(xi, -xi)`, one gradient step.  `batches`: optional explicit draws (one length-batchsize vector per epoch)
    instead of `rng` (the example uses the unseeded global RNG).  Only D = 1 chains, like the example."""
    rng = rng or np.random.default_rng()
    trafo = copy.deepcopy(initial_trafo)
    state = copy.deepcopy(optstate) if optstate is not None else setup(optimizer, trafo)
    hist = []
    n = nepochs if batches is None else len(batches)
    for i in range(n):
        b = rng.standard_normal(batchsize) if batches is None else np.asarray(batches[i])
        xi = np.concatenate([b, -b]).astype(dtype)[None, :]          # 1 x (2 batchsize): antithetic sampling
        Xd = B200Matrix.from_host(xi, ctx) if ctx is not None else B200Matrix.from_host(xi)
        nelbo, d_trafo = nELBO_trafograd(trafo, Xd, target)
        state, trafo = update(optimizer, state, trafo, d_trafo)
        hist.append(float(nelbo))
    return {"result": trafo, "optimizer_state": state, "nelbo_history": list(nelbo_history or []) + hist}
