"""Literal CPU restatement of the reference's batched trafo-chain path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line it follows (paths relative to /root/reference).  Operation
order, "ladj evaluated at y" choices, the un-normalised Householder division,
per-batch 1/N normalisation and the contiguous un-shuffled batching are kept
exactly as the reference writes them.

The formulas are written once, generic over an array namespace (numpy or
torch), so the same restatement serves as
  * the float64 / float32 / longdouble value oracle (numpy), and
  * the gradient oracle: torch-float64 reverse-mode autograd of the *literal*
    forward formulas plays the role Zygote plays in the reference
    (src/optimize_whitening.jl:18-22).

Sample matrices are D x N, one sample per column (Julia column-major D x N is
the same memory as a C-contiguous N x D numpy array; this module works on the
logical D x N view and leaves memory layout to the callers).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Sequence, Tuple, Union

import numpy as np

LOG2PI = math.log(2.0 * math.pi)  # Distributions.log2π (src/EuclidianNormalizingFlows.jl:36)


# --------------------------------------------------------------------------
# array-namespace plumbing
# --------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ns(*xs):
    for x in xs:
        if _is_torch(x):
            import torch
            return torch
    return np


def _float_dtype(*xs):
    """Julia's float(promote_type(...)) for the numeric kinds we meet:
    Ints promote to the widest float present, or Float64 if there is none
    (src/center_stretch.jl:5, src/johnson_trafo.jl:30)."""
    best = None
    for x in xs:
        dt = np.asarray(x).dtype
        if dt.kind == "f":
            best = dt if best is None else np.promote_types(best, dt)
    return np.dtype(np.float64) if best is None else best


def _prep(x, *params):
    """Promote (x, params...) to one float dtype; params broadcast along rows
    of a D x N matrix exactly like Julia's `f.(x, a, b, c)` with length-D a."""
    if _ns(x, *params) is not np:
        import torch
        x = torch.as_tensor(x, dtype=torch.float64)
        out = [x]
        for p in params:
            p = torch.as_tensor(p, dtype=torch.float64)
            if p.ndim == 1 and x.ndim == 2:
                p = p[:, None]
            out.append(p)
        return out
    dt = _float_dtype(x, *params)
    x = np.asarray(x, dtype=dt)
    out = [x]
    for p in params:
        p = np.asarray(p, dtype=dt)
        if p.ndim == 1 and x.ndim == 2:
            p = p[:, None]
        out.append(p)
    return out


# --------------------------------------------------------------------------
# scalar kernels (broadcast elementwise)
# --------------------------------------------------------------------------
def center_stretch(x, a, b, c):
    """src/center_stretch.jl:4-8."""
    x, a, b, c = _prep(x, a, b, c)
    xp = _ns(x)
    exp_abs_bx = xp.exp(xp.abs(b * x))
    om = 1 - exp_abs_bx
    return xp.sign(x) * xp.log((xp.sqrt(om * om * xp.exp(2 * b * a) + 4 * exp_abs_bx)
                                - om * xp.exp(b * a)) / 2) / b + c


def center_contract(x, a, b, c):
    """src/center_stretch.jl:11-15."""
    x, a, b, c = _prep(x, a, b, c)
    xp = _ns(x)
    u = x - c
    return (xp.log(1 + xp.exp(b * (u - a))) - xp.log(1 + xp.exp(-b * (u + a)))) / b


def center_contract_ladj(x, a, b, c):
    """src/center_stretch.jl:17-22."""
    x, a, b, c = _prep(x, a, b, c)
    xp = _ns(x)
    u = x - c
    dy_dx = 1 / (1 + xp.exp(-b * (u - a))) + 1 / (1 + xp.exp(b * (u + a)))
    return xp.log(xp.abs(dy_dx))


def johnsontrafo(x, gamma, delta, xi, lam):
    """src/johnson_trafo.jl:29-32."""
    x, gamma, delta, xi, lam = _prep(x, gamma, delta, xi, lam)
    xp = _ns(x)
    return gamma + delta * xp.asinh((x - xi) / lam)


def johnsontrafo_inv(x, gamma, delta, xi, lam):
    """src/johnson_trafo.jl:34-37."""
    x, gamma, delta, xi, lam = _prep(x, gamma, delta, xi, lam)
    xp = _ns(x)
    return lam * xp.sinh((x - gamma) / delta) + xi


def deriv_johnsontrafo(x, gamma, delta, xi, lam):
    """src/johnson_trafo.jl:39-42."""
    x, gamma, delta, xi, lam = _prep(x, gamma, delta, xi, lam)
    xp = _ns(x)
    z = (x - xi) / lam
    return (delta / lam) * (1 / xp.sqrt(1 + z * z))


def deriv_johnsontrafo_inv(x, gamma, delta, xi, lam):
    """src/johnson_trafo.jl:44-47."""
    x, gamma, delta, xi, lam = _prep(x, gamma, delta, xi, lam)
    xp = _ns(x)
    return lam * xp.cosh((x - gamma) / delta) / delta


def johnsontrafo_ladj(x, gamma, delta, xi, lam):
    """src/johnson_trafo.jl:49-52."""
    d = deriv_johnsontrafo(x, gamma, delta, xi, lam)
    xp = _ns(d)
    return xp.log(xp.abs(d))


def johnsontrafo_inv_ladj(x, gamma, delta, xi, lam):
    """src/johnson_trafo.jl:54-57 (not used by the batched path, only by
    test/test_johnson_trafo.jl:29,45-48)."""
    d = deriv_johnsontrafo_inv(x, gamma, delta, xi, lam)
    xp = _ns(d)
    return xp.log(xp.abs(d))


def std_normal_logpdf(x):
    """src/optimize_whitening.jl:4."""
    return -(x * x + LOG2PI) / 2


def sum_ladjs(ladjs):
    """src/abstract_trafo.jl:7-9.  Scalar -> itself, vector (one sample) ->
    sum, matrix D x N -> length-N row (the reference returns a 1 x N Adjoint)."""
    if ladjs.ndim == 0:
        return ladjs
    if ladjs.ndim == 1:
        return ladjs.sum()
    return ladjs.sum(0)


# --------------------------------------------------------------------------
# Householder kernels
# --------------------------------------------------------------------------
def householder_trafo(v, x):
    """src/householder_trafo.jl:4-19: k = (v'x)/(v'v); y = muladd(-2k, v, x).
    v is NOT assumed normalised."""
    xp = _ns(v, x)
    if xp is np:
        dt = np.promote_types(np.asarray(v).dtype, np.asarray(x).dtype)
        v = np.asarray(v, dtype=dt)
        x = np.asarray(x, dtype=dt)
    k = (v @ x) / (v @ v)
    if x.ndim == 2:
        return (-2 * k)[None, :] * v[:, None] + x
    return (-2 * k) * v + x


def chained_householder_trafo(V, x):
    """src/householder_trafo.jl:71-85: reflections applied for columns
    1..K of V in order (order pinned by test/test_householder_trafo.jl:40)."""
    y = x
    for i in range(V.shape[1]):
        y = householder_trafo(V[:, i], y)
    return y


def householder_trafo_pullback_v(v, x, dO):
    """src/householder_trafo.jl:22-40, literal (the non-'readable' version)."""
    xp = _ns(v, x, dO)
    inrm = 1 / xp.sqrt((v * v).sum())
    inrm_2 = inrm * inrm
    if x.ndim == 1:
        x = x[:, None]
        dO = dO[:, None]
    w_x = inrm * (v @ x)          # 1 x N
    w_dO = inrm * (v @ dO)        # 1 x N
    vv = v[:, None]
    dw_v = (-2 * vv * (w_dO[None, :] * x + w_x[None, :] * dO)).sum(0)   # 1 x N
    return (inrm * (-2 * (w_x[None, :] * dO + w_dO[None, :] * x)
                    - inrm_2 * dw_v[None, :] * vv)).sum(1)              # D (x 1)


def householder_trafo_pullback_x(v, x, dO):
    """src/householder_trafo.jl:45-54: dx = H dO."""
    return householder_trafo(v, dO)


def chained_householder_trafo_pullback_V(V, x, y, dO):
    """src/householder_trafo.jl:88-103: reverse sweep with recomputation."""
    xp = _ns(V, x, y, dO)
    dV_cols = [None] * V.shape[1]
    z = y
    D_ = dO
    for i in reversed(range(V.shape[1])):
        v = V[:, i]
        z = householder_trafo(v, z)
        dV_cols[i] = householder_trafo_pullback_v(v, z, D_)
        D_ = householder_trafo(v, D_)
    if xp is np:
        assert np.allclose(z, x)       # `@assert z ≈ x`, src/householder_trafo.jl:101
    return xp.stack(dV_cols, 1)


def chained_householder_trafo_pullback_x(V, x, y, dO):
    """src/householder_trafo.jl:105-114."""
    dx = dO
    for i in reversed(range(V.shape[1])):
        dx = householder_trafo(V[:, i], dx)
    return dx


# --------------------------------------------------------------------------
# trafo structs (same field names and order as the Julia structs)
# --------------------------------------------------------------------------
@dataclass
class CenterStretch:          # src/center_stretch.jl:25-45
    a: Any = 0.0
    b: Any = 1.0
    c: Any = 0.0
    fields = ("a", "b", "c")


@dataclass
class CenterContract:         # src/center_stretch.jl:49-69
    a: Any = 0.0
    b: Any = 1.0
    c: Any = 0.0
    fields = ("a", "b", "c")


@dataclass
class JohnsonTrafo:           # src/johnson_trafo.jl:61-82
    gamma: Any = 10.0
    delta: Any = 3.5
    xi: Any = 10.0
    lam: Any = 1.0            # Julia field name: lambda
    fields = ("gamma", "delta", "xi", "lam")


@dataclass
class JohnsonTrafoInv:        # src/johnson_trafo.jl:86-107
    gamma: Any = 10.0
    delta: Any = 3.5
    xi: Any = 10.0
    lam: Any = 1.0
    fields = ("gamma", "delta", "xi", "lam")


@dataclass
class ScaleShiftTrafo:        # src/scale_shift_trafo.jl:4-7
    a: Any = None
    b: Any = None
    fields = ("a", "b")


@dataclass
class HouseholderTrafo:       # src/householder_trafo.jl:127-129
    V: Any = None
    fields = ("V",)


@dataclass
class Composed:
    """Base.ComposedFunction: (outer ∘ inner)(x) = outer(inner(x))."""
    outer: Any = None
    inner: Any = None
    fields = ("outer", "inner")


def compose(*fs):
    """`f1 ∘ f2 ∘ ... ∘ fn` (left-assoc like Julia): fn is applied first."""
    out = fs[0]
    for f in fs[1:]:
        out = Composed(out, f)
    return out


def flatten(f) -> list:
    """Leaf trafos in application order (innermost first)."""
    if isinstance(f, Composed):
        return flatten(f.inner) + flatten(f.outer)
    return [f]


def apply(f, x):
    """The call operators: src/center_stretch.jl:37,61; src/johnson_trafo.jl:74,99;
    src/scale_shift_trafo.jl:15-16; src/householder_trafo.jl:156-157."""
    if isinstance(f, Composed):
        return apply(f.outer, apply(f.inner, x))
    if isinstance(f, CenterStretch):
        return center_stretch(x, f.a, f.b, f.c)
    if isinstance(f, CenterContract):
        return center_contract(x, f.a, f.b, f.c)
    if isinstance(f, JohnsonTrafo):
        return johnsontrafo(x, f.gamma, f.delta, f.xi, f.lam)
    if isinstance(f, JohnsonTrafoInv):
        return johnsontrafo_inv(x, f.gamma, f.delta, f.xi, f.lam)
    if isinstance(f, ScaleShiftTrafo):
        x, a, b = _prep(x, f.a, f.b)
        return x * a + b                                   # muladd.(x, a, b)
    if isinstance(f, HouseholderTrafo):
        V = f.V
        if V.ndim == 1:
            return householder_trafo(V, x)
        return chained_householder_trafo(V, x)
    raise TypeError(f)


def with_logabsdet_jacobian(f, x, *, zygote_primal: bool = False):
    """Returns (y, ladj).  ladj: scalar for one sample (x.ndim == 1), length-N
    vector for a D x N matrix (the reference's 1 x N Adjoint row).

    zygote_primal=True reproduces what the forward pass evaluates to *under
    Zygote.pullback*: rrule(similar_fill) returns zeros as the primal
    (src/abstract_trafo.jl:30-33), so ScaleShiftTrafo contributes 0 to ladj."""
    xp = _ns(x)
    if isinstance(f, Composed):
        # ChangesOfVariables 0.1, with_logabsdet_jacobian(::ComposedFunction, x)
        # (un-vendored; PARITY UNPINNED): inner first, ladjs added.
        y_i, l_i = with_logabsdet_jacobian(f.inner, x, zygote_primal=zygote_primal)
        y, l_o = with_logabsdet_jacobian(f.outer, y_i, zygote_primal=zygote_primal)
        return y, l_i + l_o
    if isinstance(f, CenterStretch):           # src/center_stretch.jl:39-43
        y = apply(f, x)
        neg = center_contract_ladj(y, f.a, f.b, f.c)
        return y, -sum_ladjs(neg)
    if isinstance(f, CenterContract):          # src/center_stretch.jl:63-67
        y = apply(f, x)
        return y, sum_ladjs(center_contract_ladj(x, f.a, f.b, f.c))
    if isinstance(f, JohnsonTrafo):            # src/johnson_trafo.jl:76-80
        y = apply(f, x)
        return y, sum_ladjs(johnsontrafo_ladj(x, f.gamma, f.delta, f.xi, f.lam))
    if isinstance(f, JohnsonTrafoInv):         # src/johnson_trafo.jl:101-105
        y = apply(f, x)
        neg = johnsontrafo_ladj(y, f.gamma, f.delta, f.xi, f.lam)
        return y, -sum_ladjs(neg)
    if isinstance(f, ScaleShiftTrafo):         # src/scale_shift_trafo.jl:18-24
        y = apply(f, x)
        a = f.a if _is_torch(f.a) else np.asarray(f.a, dtype=y.dtype if xp is np else None)
        if xp is not np:
            import torch
            a = torch.as_tensor(a, dtype=torch.float64)
        ladj = xp.log(xp.abs(a)).sum()
        if zygote_primal:
            ladj = ladj * 0
        if x.ndim == 2:
            ladj = ladj + xp.zeros(x.shape[1], dtype=y.dtype)
        return y, ladj
    if isinstance(f, HouseholderTrafo):        # src/householder_trafo.jl:159-160
        y = apply(f, x)
        if x.ndim == 2:
            return y, xp.zeros(x.shape[1], dtype=y.dtype)
        return y, xp.zeros((), dtype=y.dtype)
    raise TypeError(f)


def inverse(f):
    """src/center_stretch.jl:45,69; src/johnson_trafo.jl:82,107;
    src/scale_shift_trafo.jl:26-30; src/householder_trafo.jl:153-154;
    ComposedFunction: InverseFunctions 0.1 (un-vendored; PARITY UNPINNED):
    inverse(f ∘ g) = inverse(g) ∘ inverse(f)."""
    if isinstance(f, Composed):
        return Composed(inverse(f.inner), inverse(f.outer))
    if isinstance(f, CenterStretch):
        return CenterContract(f.a, f.b, f.c)
    if isinstance(f, CenterContract):
        return CenterStretch(f.a, f.b, f.c)
    if isinstance(f, JohnsonTrafo):
        return JohnsonTrafoInv(f.gamma, f.delta, f.xi, f.lam)
    if isinstance(f, JohnsonTrafoInv):
        return JohnsonTrafo(f.gamma, f.delta, f.xi, f.lam)
    if isinstance(f, ScaleShiftTrafo):
        a = np.asarray(f.a)
        a_inv = 1 / a
        b_inv = -a_inv * np.asarray(f.b)
        return ScaleShiftTrafo(a_inv, b_inv)
    if isinstance(f, HouseholderTrafo):
        V = np.asarray(f.V)
        return f if V.ndim == 1 else HouseholderTrafo(V[:, ::-1].copy())
    raise TypeError(f)


# --------------------------------------------------------------------------
# loss and gradient (src/optimize_whitening.jl:4-22)
# --------------------------------------------------------------------------
def mvnormal_negll_trafo(trafo, X, *, zygote_primal: bool = False):
    """src/optimize_whitening.jl:7-15."""
    nsamples = X.shape[1]
    Y, ladj = with_logabsdet_jacobian(trafo, X, zygote_primal=zygote_primal)
    ll = (std_normal_logpdf(Y).sum() + ladj.sum()) / nsamples
    return -ll


def _to_torch_params(f):
    import torch
    if isinstance(f, Composed):
        return Composed(_to_torch_params(f.outer), _to_torch_params(f.inner))
    kw = {}
    for name in f.fields:
        kw[name] = torch.tensor(np.asarray(getattr(f, name), dtype=np.float64),
                                dtype=torch.float64, requires_grad=True)
    return type(f)(**kw)


def _grads_of(ft) -> Dict[str, Any]:
    if isinstance(ft, Composed):
        return {"outer": _grads_of(ft.outer), "inner": _grads_of(ft.inner)}
    out = {}
    for name in ft.fields:
        g = getattr(ft, name).grad
        out[name] = None if g is None else g.detach().numpy().copy()
    return out


def mvnormal_negll_trafograd(trafo, X, *, zygote_primal: bool = True):
    """src/optimize_whitening.jl:18-22.  Reverse-mode AD (torch float64
    autograd standing in for Zygote, PARITY UNPINNED for the AD engine; the
    result is mathematically determined).  Returns (negll, d_trafo) with
    d_trafo a nested dict mirroring Zygote's NamedTuple:
    {'outer':..., 'inner':...} for ComposedFunction, field-name keys for leaves.

    zygote_primal=True (default, what the reference returns): the ScaleShift
    ladj value is dropped from negll (src/abstract_trafo.jl:32) while its
    gradient is kept."""
    import torch
    ft = _to_torch_params(trafo)
    Xt = torch.tensor(np.asarray(X, dtype=np.float64), dtype=torch.float64)
    negll_true = mvnormal_negll_trafo(ft, Xt, zygote_primal=False)
    negll_true.backward()
    grads = _grads_of(ft)
    if zygote_primal:
        with torch.no_grad():
            val = float(mvnormal_negll_trafo(ft, Xt, zygote_primal=True))
    else:
        val = float(negll_true.detach())
    return val, grads


# --------------------------------------------------------------------------
# optimizer and fit loop (src/optimize_whitening.jl:25-45)
# --------------------------------------------------------------------------
def _ht_normalize(V):
    """src/householder_trafo.jl:135-140."""
    V = np.asarray(V)
    if V.ndim == 1:
        return V / np.sqrt((V * V).sum())
    return V / np.sqrt((V * V).sum(0))[None, :]


@dataclass
class ADAGrad:
    """Optimisers.jl 0.2 `ADAGrad(η = 1f-1, ϵ = eps(typeof(η)))` (un-vendored;
    PARITY UNPINNED): state initialised to ϵ; acc += g²; x -= η g /(√acc + ϵ)."""
    eta: float = float(np.float32(0.1))
    epsilon: float = float(np.finfo(np.float32).eps)


def optim_setup(opt: ADAGrad, trafo):
    if isinstance(trafo, Composed):
        return {"outer": optim_setup(opt, trafo.outer), "inner": optim_setup(opt, trafo.inner)}
    return {n: np.full_like(np.asarray(getattr(trafo, n), dtype=np.float64), opt.epsilon)
            for n in trafo.fields}


def optim_update(opt: ADAGrad, state, trafo, grads):
    """Optimisers.update(state, trafo, d_trafo) (src/optimize_whitening.jl:40).
    Rebuilds the struct tree through Functors; HouseholderTrafo's functor
    re-normalises every column of V (src/householder_trafo.jl:134-146)."""
    if isinstance(trafo, Composed):
        so, to = optim_update(opt, state["outer"], trafo.outer, grads["outer"])
        si, ti = optim_update(opt, state["inner"], trafo.inner, grads["inner"])
        return {"outer": so, "inner": si}, Composed(to, ti)
    new_state, kw = {}, {}
    for n in trafo.fields:
        x = np.asarray(getattr(trafo, n), dtype=np.float64)
        g = grads[n]
        if g is None:
            new_state[n], kw[n] = state[n], x
            continue
        g = np.asarray(g, dtype=np.float64).reshape(x.shape)
        acc = state[n] + g * g
        kw[n] = x - g * opt.eta / (np.sqrt(acc) + opt.epsilon)
        new_state[n] = acc
    new = type(trafo)(**kw)
    if isinstance(new, HouseholderTrafo):
        new = HouseholderTrafo(_ht_normalize(new.V))
    return new_state, new


def batch_ranges(nsamples: int, nbatches: int) -> List[Tuple[int, int]]:
    """src/optimize_whitening.jl:31-32: batchsize = round(Int, N/nbatches)
    (ties to even, like Python's round); Iterators.partition -> contiguous
    column ranges in fixed order, last one possibly short."""
    batchsize = int(round(nsamples / nbatches))
    return [(s, min(s + batchsize, nsamples)) for s in range(0, nsamples, batchsize)]


def optimize_whitening(X, initial_trafo, optimizer: ADAGrad, *, nbatches=100, nepochs=100,
                       optstate=None, negll_history=None):
    """src/optimize_whitening.jl:25-45.  X: D x N matrix (the reference takes
    nestedview(X); flatview(batch) is a D x batchsize column range)."""
    import copy
    trafo = copy.deepcopy(initial_trafo)
    state = copy.deepcopy(optstate) if optstate is not None else optim_setup(optimizer, trafo)
    hist: List[float] = []
    for _ in range(nepochs):
        for (s, e) in batch_ranges(X.shape[1], nbatches):
            negll, d_trafo = mvnormal_negll_trafograd(trafo, X[:, s:e])
            state, trafo = optim_update(optimizer, state, trafo, d_trafo)
            hist.append(float(negll))
    return {"result": trafo, "optimizer_state": state,
            "negll_history": list(negll_history or []) + hist}


# --------------------------------------------------------------------------
# JohnsonSU distribution object and the variational (ELBO) objective
# --------------------------------------------------------------------------
@dataclass
class JohnsonSU:
    """src/johnson_trafo.jl:1-26,109-129 (`JohnsonSU <: Distribution{Univariate,Continuous}`), scalar parameters.
    Distributions.jl / StatsFuns.jl (un-vendored; PARITY UNPINNED) supply the standard-normal pdf / cdf / logcdf /
    quantile: restated with scipy.special.ndtr / log_ndtr / ndtri."""
    gamma: float = 10.0
    delta: float = 3.5
    xi: float = 10.0
    lam: float = 1.0           # Julia field name: lambda

    def params(self):          # StatsBase.params, :18
        return (self.gamma, self.delta, self.xi, self.lam)

    def mean(self):            # :24
        return self.xi - self.lam * math.exp(self.delta ** -2 / 2) * math.sinh(self.gamma / self.delta)

    def median(self):          # :25
        return self.xi + self.lam * math.sinh(-self.gamma / self.delta)

    def var(self):             # :26
        return (self.lam ** 2) / 2 * (math.exp(self.delta ** -2) - 1) * (math.exp(self.delta ** -2) * math.cosh(2 * self.gamma / self.delta) + 1)

    def _z(self, x):
        return johnsontrafo(x, self.gamma, self.delta, self.xi, self.lam)

    def pdf(self, x):          # :120
        z = self._z(x)
        return deriv_johnsontrafo(x, self.gamma, self.delta, self.xi, self.lam) * np.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)

    def logpdf(self, x):       # :123 (log of the product; evaluated as the sum of logs where the product underflows)
        z = self._z(x)
        return np.log(np.abs(deriv_johnsontrafo(x, self.gamma, self.delta, self.xi, self.lam))) - 0.5 * z * z - 0.5 * LOG2PI

    def cdf(self, x):          # :121
        from scipy.special import ndtr
        return ndtr(self._z(x))

    def logcdf(self, x):       # :124
        from scipy.special import log_ndtr
        return log_ndtr(self._z(x))

    def ccdf(self, x):         # :125
        return 1 - self.cdf(x)

    def logccdf(self, x):      # :126
        return np.log(1 - self.cdf(x))

    def quantile(self, p):     # :129
        from scipy.special import ndtri
        return johnsontrafo_inv(ndtri(np.asarray(p, dtype=np.float64)), self.gamma, self.delta, self.xi, self.lam)

    def rand(self, rng, n):
        """Distributions' default sampler for a continuous univariate distribution without its own `rand`:
        quantile(d, rand()) (inverse-cdf sampling)."""
        return self.quantile(rng.uniform(size=n))


@dataclass
class GaussMixture:
    """Element-wise target log-density of examples/nf_variational_1d.jl:25-27:
    my_ll(x) = log(0.3 N(x-2) + 0.5 N(x-5) + 0.2 N(x+1)), generalised to K components with widths."""
    weights: Any = (0.3, 0.5, 0.2)
    means: Any = (2.0, 5.0, -1.0)
    sigmas: Any = (1.0, 1.0, 1.0)

    def logpdf(self, x):
        xp = _ns(x)
        acc = None
        for w, mu, sg in zip(self.weights, self.means, self.sigmas):
            t = (x - mu) / sg
            term = (w / sg) * xp.exp(-(t * t + LOG2PI) / 2)       # w * std_normal_pdf, :23
            acc = term if acc is None else acc + term
        return xp.log(acc)


def nELBO(trafo, xi, target=None):
    """examples/nf_variational_1d.jl:29-41, literal: `xi` is what the example passes, a (N_samps x xi_dim) matrix
    (there: (2 batchsize) x 1) that with_logabsdet_jacobian treats like any D x N matrix -- its ROWS are the draws,
    the length-1 parameter vectors broadcast along them, and `ladj` is summed over them."""
    target = target or GaussMixture()
    z, ladj = with_logabsdet_jacobian(trafo, xi)
    xi_dim = xi.shape[1]
    n_samps = xi.shape[0]
    elbo = (target.logpdf(z).sum() + ladj.sum()) / n_samps - 0.5 * (LOG2PI + 1) * xi_dim
    return -elbo


def nELBO_trafograd(trafo, xi, target=None):
    """examples/nf_variational_1d.jl:43-47 (torch float64 autograd in the role of Zygote)."""
    import torch
    ft = _to_torch_params(trafo)
    val = nELBO(ft, torch.tensor(np.asarray(xi, dtype=np.float64), dtype=torch.float64), target)
    val.backward()
    return float(val.detach()), _grads_of(ft)


def optimise_ELBO(initial_trafo, optimizer: ADAGrad, batches, *, target=None, optstate=None, nelbo_history=None):
    """examples/nf_variational_1d.jl:49-69 with the standard-normal draws of every epoch passed in (`batches`: list of
    length-batchsize vectors; the example draws them from the unseeded global RNG): antithetic pairs `vcat(xi, -xi)`."""
    import copy
    trafo = copy.deepcopy(initial_trafo)
    state = copy.deepcopy(optstate) if optstate is not None else optim_setup(optimizer, trafo)
    hist = []
    for b in batches:
        b = np.asarray(b, dtype=np.float64).reshape(-1, 1)
        xi = np.vstack([b, -b])                                    # antithetic sampling, :62
        nelbo, d_trafo = nELBO_trafograd(trafo, xi, target)
        state, trafo = optim_update(optimizer, state, trafo, d_trafo)
        hist.append(nelbo)
    return {"result": trafo, "optimizer_state": state, "nelbo_history": list(nelbo_history or []) + hist}
