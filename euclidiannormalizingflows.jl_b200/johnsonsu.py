"""JohnsonSU distribution object (src/johnson_trafo.jl:1-26,109-129) over the batched device kernels.

The scalar summaries (mean, median, var, params) are the reference's closed forms; every density / cdf / quantile
evaluation of an array is one call into libenf_b200.so (enf_johnsonsu, csrc/enf_johnsonsu.cu).  Arrays may be
B200Matrix (stay on the device) or numpy (uploaded, evaluated, downloaded); there is no numpy fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L
from .device import B200Matrix, default_context, enf_dtype


class JohnsonSU:
    """JohnsonSU(; gamma = 10, delta = 3.5, xi = 10, lambda = 1): src/johnson_trafo.jl:1-12 (`lam` is Julia's `lambda`)."""

    def __init__(self, gamma=10.0, delta=3.5, xi=10.0, lam=1.0):
        self.gamma, self.delta, self.xi, self.lam = float(gamma), float(delta), float(xi), float(lam)

    # ---- scalar summaries (src/johnson_trafo.jl:15-26)
    def minimum(self):
        return -math.inf

    def maximum(self):
        return math.inf

    def params(self):
        return (self.gamma, self.delta, self.xi, self.lam)

    def mean(self):
        return self.xi - self.lam * math.exp(self.delta ** -2 / 2) * math.sinh(self.gamma / self.delta)

    def median(self):
        return self.xi + self.lam * math.sinh(-self.gamma / self.delta)

    def var(self):
        e = math.exp(self.delta ** -2)
        return (self.lam ** 2) / 2 * (e - 1) * (e * math.cosh(2 * self.gamma / self.delta) + 1)

    location = mean      # Distributions.location(d) = mean(d), :21
    scale = var          # Distributions.scale(d) = var(d), :22

    # ---- batched evaluations
    def _eval(self, op: int, x, ctx=None):
        p = (C.c_double * 4)(*self.params())
        if isinstance(x, B200Matrix):
            out = x.empty_like()
            L.check(x.ctx._lib.enf_johnsonsu(x.ctx.handle, enf_dtype(x.dtype), op, p, C.c_void_p(x.ptr), x.D * x.N, C.c_void_p(out.ptr)),
                    x.ctx.handle)
            return out
        a = np.asarray(x)
        dt = np.dtype(np.float32) if a.dtype == np.float32 else np.dtype(np.float64)
        flat = np.ascontiguousarray(a, dtype=dt).reshape(1, -1)
        d = B200Matrix.from_host(flat, ctx or default_context())
        L.check(d.ctx._lib.enf_johnsonsu(d.ctx.handle, enf_dtype(dt), op, p, C.c_void_p(d.ptr), d.N, C.c_void_p(d.ptr)), d.ctx.handle)
        res = d.to_host().reshape(a.shape)
        return res[()] if a.ndim == 0 else res

    def pdf(self, x, ctx=None):
        return self._eval(L.ENF_JSU_PDF, x, ctx)

    def logpdf(self, x, ctx=None):
        return self._eval(L.ENF_JSU_LOGPDF, x, ctx)

    def cdf(self, x, ctx=None):
        return self._eval(L.ENF_JSU_CDF, x, ctx)

    def logcdf(self, x, ctx=None):
        return self._eval(L.ENF_JSU_LOGCDF, x, ctx)

    def ccdf(self, x, ctx=None):
        return self._eval(L.ENF_JSU_CCDF, x, ctx)

    def logccdf(self, x, ctx=None):
        return self._eval(L.ENF_JSU_LOGCCDF, x, ctx)

    def quantile(self, p, ctx=None):
        return self._eval(L.ENF_JSU_QUANTILE, p, ctx)

    def rand(self, rng: np.random.Generator, n: int, dtype=np.float64, ctx=None):
        """n draws by inverse-cdf sampling, quantile(d, rand()): what Distributions.jl does for a continuous univariate
        distribution that defines no sampler of its own (test/test_johnson_trafo.jl:12-16 compares exactly this with
        johnsontrafo_inv.(randn))."""
        return self.quantile(rng.uniform(size=n).astype(dtype), ctx)

    def __repr__(self):
        return f"JohnsonSU(gamma={self.gamma}, delta={self.delta}, xi={self.xi}, lam={self.lam})"
