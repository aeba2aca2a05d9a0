"""Host-side mirror of the reference's trafo API for the accelerated path.

Same names, argument meaning and return shapes as the reference's generic
functions (paths relative to the reference repository):

  CenterStretch / CenterContract     src/center_stretch.jl:25-69
  JohnsonTrafo / JohnsonTrafoInv     src/johnson_trafo.jl:61-107
  ScaleShiftTrafo                    src/scale_shift_trafo.jl:4-30
  HouseholderTrafo                   src/householder_trafo.jl:127-160
  `∘` (Base.ComposedFunction)        ComposedFunction / compose()
  inverse                            InverseFunctions.inverse methods of the above
  with_logabsdet_jacobian            ChangesOfVariables methods of the above
  mvnormal_negll_trafo{,grad}        src/optimize_whitening.jl:7-22

Nothing numeric happens here: a trafo tree is flattened (innermost first) into
the op list of an `enf_chain` and every evaluation is one call into
libenf_b200.so.  `inverse` is a struct rewrite, exactly as in the reference.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

from . import _lib as L
from .device import B200Matrix, Context, default_context, enf_dtype


def _vec(p) -> np.ndarray:
    return np.atleast_1d(np.asarray(p))


class Trafo:
    """Base of the callable trafo structs (`<: Function` in the reference)."""
    fields: Tuple[str, ...] = ()
    kind: int = -1

    def __call__(self, x):
        return _evaluate(self, x, want_ladj=False)[0]

    def __matmul__(self, other):
        """`f @ g` is the reference's `f ∘ g`."""
        return ComposedFunction(self, other)

    def __eq__(self, other):
        return type(self) is type(other) and all(
            np.array_equal(np.asarray(getattr(self, n)), np.asarray(getattr(other, n))) for n in self.fields)

    def __hash__(self):
        return hash((type(self).__name__,) + tuple(np.asarray(getattr(self, n)).tobytes() for n in self.fields))

    def __repr__(self):
        return f"{type(self).__name__}(" + ", ".join(f"{n}={getattr(self, n)!r}" for n in self.fields) + ")"


class CenterStretch(Trafo):
    fields = ("a", "b", "c")
    kind = L.ENF_CENTER_STRETCH

    def __init__(self, a=0.0, b=1.0, c=0.0):
        self.a, self.b, self.c = a, b, c


class CenterContract(Trafo):
    fields = ("a", "b", "c")
    kind = L.ENF_CENTER_CONTRACT

    def __init__(self, a=0.0, b=1.0, c=0.0):
        self.a, self.b, self.c = a, b, c


class JohnsonTrafo(Trafo):
    fields = ("gamma", "delta", "xi", "lam")   # Julia field name of `lam`: lambda
    kind = L.ENF_JOHNSON

    def __init__(self, gamma=10.0, delta=3.5, xi=10.0, lam=1.0):
        self.gamma, self.delta, self.xi, self.lam = gamma, delta, xi, lam


class JohnsonTrafoInv(Trafo):
    fields = ("gamma", "delta", "xi", "lam")
    kind = L.ENF_JOHNSON_INV

    def __init__(self, gamma=10.0, delta=3.5, xi=10.0, lam=1.0):
        self.gamma, self.delta, self.xi, self.lam = gamma, delta, xi, lam


class ScaleShiftTrafo(Trafo):
    fields = ("a", "b")
    kind = L.ENF_SCALE_SHIFT

    def __init__(self, a, b):
        self.a, self.b = a, b


class HouseholderTrafo(Trafo):
    fields = ("V",)
    kind = L.ENF_HOUSEHOLDER

    def __init__(self, V):
        self.V = np.asarray(V)


class ComposedFunction(Trafo):
    """Base.ComposedFunction: (outer ∘ inner)(x) = outer(inner(x))."""
    fields = ("outer", "inner")

    def __init__(self, outer, inner):
        self.outer, self.inner = outer, inner

    def __eq__(self, other):
        return isinstance(other, ComposedFunction) and self.outer == other.outer and self.inner == other.inner

    def __hash__(self):
        return hash((hash(self.outer), hash(self.inner)))


def compose(*fs) -> Trafo:
    """`f1 ∘ f2 ∘ ... ∘ fn` (left-associative like Julia): fn is applied first."""
    out = fs[0]
    for f in fs[1:]:
        out = ComposedFunction(out, f)
    return out


def flatten(f) -> List[Trafo]:
    """Leaf trafos in application order (innermost first): the order
    ChangesOfVariables evaluates a ComposedFunction in."""
    if isinstance(f, ComposedFunction):
        return flatten(f.inner) + flatten(f.outer)
    return [f]


def inverse(f):
    """InverseFunctions.inverse: src/center_stretch.jl:45,69; src/johnson_trafo.jl:82,107;
    src/scale_shift_trafo.jl:26-30; src/householder_trafo.jl:153-154; and
    inverse(f ∘ g) = inverse(g) ∘ inverse(f).  O(params) on the host, no data touched."""
    if isinstance(f, ComposedFunction):
        return ComposedFunction(inverse(f.inner), inverse(f.outer))
    if isinstance(f, CenterStretch):
        return CenterContract(f.a, f.b, f.c)
    if isinstance(f, CenterContract):
        return CenterStretch(f.a, f.b, f.c)
    if isinstance(f, JohnsonTrafo):
        return JohnsonTrafoInv(f.gamma, f.delta, f.xi, f.lam)
    if isinstance(f, JohnsonTrafoInv):
        return JohnsonTrafo(f.gamma, f.delta, f.xi, f.lam)
    if isinstance(f, ScaleShiftTrafo):
        a_inv = 1.0 / np.asarray(f.a, dtype=np.result_type(np.asarray(f.a).dtype, np.float32))
        return ScaleShiftTrafo(a_inv, -a_inv * np.asarray(f.b))
    if isinstance(f, HouseholderTrafo):
        return f if f.V.ndim == 1 else HouseholderTrafo(f.V[:, ::-1].copy())
    raise TypeError(f"not a trafo: {f!r}")


# ------------------------------------------------------------------ chain plumbing
def _param_dtype(leaves) -> np.dtype:
    dt = None
    for f in leaves:
        for n in f.fields:
            a = np.asarray(getattr(f, n))
            if a.dtype.kind == "f":
                dt = a.dtype if dt is None else np.promote_types(dt, a.dtype)
    return np.dtype(np.float64) if dt is None else dt


def result_dtype(f, x_dtype) -> np.dtype:
    """float(promote_type(eltype(x), eltype(params)...)) (src/center_stretch.jl:5)."""
    dt = np.promote_types(_param_dtype(flatten(f)), np.dtype(x_dtype))
    return np.dtype(np.float32) if dt == np.float32 else np.dtype(np.float64)


def pack_params(leaves, D: int, dtype) -> np.ndarray:
    """Packed parameter vector in struct-field order; scalar fields are
    expanded to length D (the C ABI takes vectors only)."""
    out = []
    for f in leaves:
        if isinstance(f, HouseholderTrafo):
            V = np.asarray(f.V, dtype=dtype)
            V = V[:, None] if V.ndim == 1 else V
            if V.shape[0] != D:
                raise ValueError(f"HouseholderTrafo has {V.shape[0]} rows, samples have {D}")
            out.append(np.asfortranarray(V).ravel(order="F"))
        else:
            for n in f.fields:
                p = np.asarray(getattr(f, n), dtype=dtype)
                if p.ndim == 0:
                    p = np.full(D, p, dtype=dtype)
                if p.shape != (D,):
                    raise ValueError(f"{type(f).__name__}.{n} has shape {p.shape}, samples have {D} rows")
                out.append(p)
    return np.ascontiguousarray(np.concatenate(out))


def _signature(leaves) -> tuple:
    return tuple((f.kind, (1 if f.V.ndim == 1 else f.V.shape[1]) if isinstance(f, HouseholderTrafo) else 0) for f in leaves)


class Chain:
    """An enf_chain handle cached on the context per (dtype, D, structure)."""

    def __init__(self, ctx: Context, handle, nparams: int, dtype):
        self.ctx, self.handle, self.nparams, self.dtype = ctx, handle, nparams, np.dtype(dtype)
        self.last: Optional[np.ndarray] = None

    def describe(self) -> str:
        buf = C.create_string_buffer(512)
        L.check(self.ctx._lib.enf_chain_describe(self.handle, buf, 512), self.ctx.handle)
        return buf.value.decode()


def get_chain(f, D: int, dtype, ctx: Optional[Context] = None) -> Chain:
    ctx = ctx or default_context()
    dtype = np.dtype(dtype)
    leaves = flatten(f)
    for lf in leaves:
        if lf.kind < 0:
            raise TypeError(f"not a trafo: {lf!r}")
    key = (dtype.str, D, _signature(leaves))
    packed = pack_params(leaves, D, dtype)
    ent = ctx._chains.get(key)
    if ent is None:
        ops = (L.enf_op * len(leaves))()
        off = 0
        for i, lf in enumerate(leaves):
            K = key[2][i][1]
            n = D * K if isinstance(lf, HouseholderTrafo) else D * len(lf.fields)
            ops[i].kind, ops[i].K = lf.kind, K
            ops[i].params = packed.ctypes.data + off * dtype.itemsize
            off += n
        h = C.c_void_p()
        L.check(ctx._lib.enf_chain_create(ctx.handle, enf_dtype(dtype), D, len(leaves), ops, C.byref(h)), ctx.handle)
        ch = Chain(ctx, h, packed.size, dtype)
        ch.last = packed
        ctx._chains[key] = {"handle": h, "chain": ch}
        return ch
    ch = ent["chain"]
    if ch.last is None or not np.array_equal(ch.last, packed):
        L.check(ctx._lib.enf_chain_set_params(ch.handle, packed.ctypes.data_as(C.c_void_p)), ctx.handle)
        ch.last = packed
    return ch


# ------------------------------------------------------------------ evaluation
def _evaluate(f, x, want_ladj: bool, out=None, ctx=None):
    """(y, ladj) for a device matrix, a host matrix (through the host-buffer
    pipeline of enf_forward_ladj_host) or a single host sample vector."""
    if isinstance(x, B200Matrix):
        dt = result_dtype(f, x.dtype)
        if dt != x.dtype:
            # float(promote_type(eltype(x), eltype(params)...)) (src/center_stretch.jl:5): Float32 samples meeting Float64
            # parameters give a Float64 result.  A chain is all-f32 or all-f64, so the samples are widened on the device.
            if out is not None:
                raise TypeError(f"device samples are {x.dtype} but the chain promotes to {dt}: out= buffers need converted samples")
            x = x.astype(dt)
        ch = get_chain(f, x.D, dt, x.ctx)
        y = out[0] if out is not None else x.empty_like()
        lib = x.ctx._lib
        if want_ladj:
            ladj = out[1] if out is not None else B200Matrix(x.ctx, 1, x.N, dt)
            L.check(lib.enf_forward_ladj(ch.handle, C.c_void_p(x.ptr), x.N, C.c_void_p(y.ptr), C.c_void_p(ladj.ptr)), x.ctx.handle)
            return y, ladj
        L.check(lib.enf_forward(ch.handle, C.c_void_p(x.ptr), x.N, C.c_void_p(y.ptr)), x.ctx.handle)
        return y, None
    x = np.asarray(x)
    single = x.ndim == 1
    X = x[:, None] if single else x
    if X.ndim != 2:
        raise ValueError("expected a sample vector or a D x N sample matrix")
    dt = result_dtype(f, X.dtype if X.dtype.kind == "f" else np.float64)
    ctx = ctx or default_context()
    ch = get_chain(f, X.shape[0], dt, ctx)
    Xf = np.asfortranarray(X, dtype=dt)
    if out is not None:
        Y, ladj = out
        if not (Y.flags.f_contiguous and Y.dtype == dt and Y.shape == Xf.shape):
            raise ValueError("out[0] must be a column-major array of the result dtype and shape")
    else:
        Y = np.empty_like(Xf, order="F")
        ladj = np.empty((1, X.shape[1]), dtype=dt) if want_ladj else None
    L.check(ctx._lib.enf_forward_ladj_host(ch.handle, Xf.ctypes.data_as(C.c_void_p), X.shape[1],
                                           Y.ctypes.data_as(C.c_void_p),
                                           ladj.ctypes.data_as(C.c_void_p) if want_ladj else None), ctx.handle)
    if single:
        return Y[:, 0], (ladj[0, 0] if want_ladj else None)
    return Y, ladj


def with_logabsdet_jacobian(f, x, out=None, ctx=None):
    """ChangesOfVariables.with_logabsdet_jacobian(f, x) -> (y, ladj).
    ladj is a 1 x N row for a D x N matrix (the reference's `Adjoint` row,
    src/abstract_trafo.jl:9) and a scalar for a single sample vector.
    out=(y, ladj): optional preallocated results (B200Matrix pair for device
    input; column-major numpy arrays, e.g. pinned ones, for host input).
    ctx: device context for host input (default: device 0)."""
    return _evaluate(f, x, want_ladj=True, out=out, ctx=ctx)


def mvnormal_negll_trafo(trafo, X) -> float:
    """src/optimize_whitening.jl:7-15 (true value, ScaleShift ladj included)."""
    X = _as_device(X, trafo)
    ch = get_chain(trafo, X.D, X.dtype, X.ctx)
    out = C.c_double()
    L.check(X.ctx._lib.enf_negll(ch.handle, C.c_void_p(X.ptr), X.N, C.byref(out)), X.ctx.handle)
    return out.value


def _as_device(X, trafo) -> B200Matrix:
    if isinstance(X, B200Matrix):
        return X.astype(result_dtype(trafo, X.dtype))      # promotion as in src/center_stretch.jl:5
    X = np.asarray(X)
    dt = result_dtype(trafo, X.dtype if X.dtype.kind == "f" else np.float64)
    return B200Matrix.from_host(np.asarray(X, dtype=dt))


def unpack_grads(trafo, flat: np.ndarray, D: int):
    """Packed gradient vector -> the nested structure Zygote returns:
    {'outer':…, 'inner':…} for ComposedFunction, field-name keys for leaves
    (SURVEY §3.4).  HouseholderTrafo with a vector V gets a D x 1 matrix like
    the reference (src/householder_trafo.jl:39)."""
    pos = [0]

    def rec(f):
        if isinstance(f, ComposedFunction):
            inner = rec(f.inner)      # application order: inner first
            outer = rec(f.outer)
            return {"outer": outer, "inner": inner}
        out = {}
        if isinstance(f, HouseholderTrafo):
            K = 1 if f.V.ndim == 1 else f.V.shape[1]
            out["V"] = flat[pos[0]:pos[0] + D * K].reshape((D, K), order="F").copy()
            pos[0] += D * K
            return out
        for n in f.fields:
            g = flat[pos[0]:pos[0] + D].copy()
            pos[0] += D
            out[n] = g if np.ndim(getattr(f, n)) else g.sum()   # scalar field: un-broadcast (Zygote sums)
        return out

    return rec(trafo)


def mvnormal_negll_trafograd(trafo, X, *, zygote_primal: bool = True, group: bool = False):
    """src/optimize_whitening.jl:18-22 -> (negll, d_trafo).

    zygote_primal=True (default) returns the loss value the reference returns:
    under Zygote the ScaleShiftTrafo ladj is evaluated as zeros
    (src/abstract_trafo.jl:30-33), the gradient is unaffected.
    group=True: X holds this rank's columns of a batch sharded over the NCCL
    group of the context; loss and gradients are those of the whole batch."""
    X = _as_device(X, trafo)
    ch = get_chain(trafo, X.D, X.dtype, X.ctx)
    negll = C.c_double()
    g = np.empty(ch.nparams, dtype=X.dtype)
    flags = L.ENF_NEGLL_ZYGOTE_PRIMAL if zygote_primal else 0
    fn = X.ctx._lib.enf_negll_grad_group if group else X.ctx._lib.enf_negll_grad
    L.check(fn(ch.handle, C.c_void_p(X.ptr), X.N, flags, C.byref(negll), g.ctypes.data_as(C.c_void_p)), X.ctx.handle)
    return negll.value, unpack_grads(trafo, g, X.D)
