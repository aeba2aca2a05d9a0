"""optimize_whitening (src/optimize_whitening.jl:25-45) on device-resident samples.

The fit loop stays on the host exactly as in the reference (L3 of SURVEY §1):
per batch one fused loss+gradient call into the library, then the optimizer
update and the struct rebuild.  Optimisers.jl is an un-vendored dependency of
the reference; `ADAGrad` restates its published 0.2 rule (PARITY UNPINNED,
SURVEY §8c), including the re-normalisation of HouseholderTrafo columns that
the reference's functor performs on every rebuild (src/householder_trafo.jl:134-146).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .device import B200Matrix, Context, default_context
from .trafos import (ComposedFunction, HouseholderTrafo, Trafo, flatten, mvnormal_negll_trafograd, result_dtype)


@dataclass
class ADAGrad:
    """Optimisers.ADAGrad(η = 1f-1, ϵ = eps(typeof(η))): state starts at ϵ,
    acc += g², x -= η g / (√acc + ϵ)."""
    eta: float = float(np.float32(0.1))
    epsilon: float = float(np.finfo(np.float32).eps)


def setup(opt: ADAGrad, trafo):
    """Optimisers.setup(optimizer, trafo)."""
    if isinstance(trafo, ComposedFunction):
        return {"outer": setup(opt, trafo.outer), "inner": setup(opt, trafo.inner)}
    return {n: np.full_like(np.asarray(getattr(trafo, n), dtype=np.float64), opt.epsilon) for n in trafo.fields}


def _ht_normalize(V):
    V = np.asarray(V)
    if V.ndim == 1:
        return V / np.sqrt((V * V).sum())
    return V / np.sqrt((V * V).sum(0))[None, :]


def update(opt: ADAGrad, state, trafo, grads):
    """Optimisers.update(state, trafo, d_trafo) -> (state, trafo)."""
    if isinstance(trafo, ComposedFunction):
        so, to = update(opt, state["outer"], trafo.outer, grads["outer"])
        si, ti = update(opt, state["inner"], trafo.inner, grads["inner"])
        return {"outer": so, "inner": si}, ComposedFunction(to, ti)
    new_state, kw = {}, {}
    for n in trafo.fields:
        x = np.asarray(getattr(trafo, n))
        g = np.asarray(grads[n], dtype=np.float64).reshape(x.shape)
        acc = state[n] + g * g
        kw[n] = (x - g * opt.eta / (np.sqrt(acc) + opt.epsilon)).astype(x.dtype if x.dtype.kind == "f" else np.float64)
        new_state[n] = acc
    new = type(trafo)(**kw)
    if isinstance(new, HouseholderTrafo):
        new = HouseholderTrafo(_ht_normalize(new.V))
    return new_state, new


def batch_ranges(nsamples: int, nbatches: int) -> List[Tuple[int, int]]:
    """src/optimize_whitening.jl:31-32: batchsize = round(Int, N / nbatches)
    (ties to even); contiguous column ranges in fixed order, no shuffling, the
    last one possibly short."""
    batchsize = int(round(nsamples / nbatches))
    if batchsize < 1:
        raise ValueError("nbatches exceeds the number of samples")
    return [(s, min(s + batchsize, nsamples)) for s in range(0, nsamples, batchsize)]


def _pack_tree(trafo, tree, D):
    """nested per-field structure (optimizer state) -> packed float64 vector in parameter order."""
    out = []
    for lf, st in zip(flatten(trafo), _leaves_of(trafo, tree)):
        for n in lf.fields:
            a = np.asarray(st[n], dtype=np.float64)
            if a.ndim == 0 and D > 1:
                raise TypeError("the device loop needs vector-valued parameters (one value per row)")
            out.append(np.asfortranarray(a.reshape(D, -1)).ravel(order="F"))
    return np.ascontiguousarray(np.concatenate(out))


def _leaves_of(trafo, tree):
    if isinstance(trafo, ComposedFunction):
        return _leaves_of(trafo.inner, tree["inner"]) + _leaves_of(trafo.outer, tree["outer"])
    return [tree]


def _unpack_tree(trafo, flat, D):
    pos = [0]

    def rec(f):
        if isinstance(f, ComposedFunction):
            inner = rec(f.inner)
            outer = rec(f.outer)
            return {"outer": outer, "inner": inner}
        out = {}
        for n in f.fields:
            shape = np.shape(getattr(f, n))
            k = int(np.prod(shape))
            out[n] = flat[pos[0]:pos[0] + k].reshape(shape, order="F").copy()
            pos[0] += k
        return out

    return rec(trafo)


def _rebuild(trafo, tree):
    if isinstance(trafo, ComposedFunction):
        return ComposedFunction(_rebuild(trafo.outer, tree["outer"]), _rebuild(trafo.inner, tree["inner"]))
    return type(trafo)(**{n: tree[n] for n in trafo.fields})


def _optimize_whitening_device(smpls, trafo, optimizer, nbatches, nepochs, optstate, group, batch_counts=None):
    """enf_optimize_whitening: the whole loop on the device."""
    import ctypes as C
    from . import _lib as L
    from .trafos import get_chain
    ctx = smpls.ctx
    # The C ABI trains one parameter per row and field.  A scalar-valued field of the reference structs is ONE shared
    # parameter (its gradient is the sum over the rows): for D > 1 that is a different model, so it is rejected here
    # instead of being expanded silently; for D = 1 the two coincide.
    if smpls.D > 1:
        for lf in flatten(trafo):
            for n in lf.fields:
                if np.ndim(getattr(lf, n)) == 0:
                    raise TypeError(f"{type(lf).__name__}.{n} is a scalar: the device loop needs vector-valued parameters "
                                    "(one value per row); use the host loop (device_loop=False) for shared scalar fields")
    ch = get_chain(trafo, smpls.D, smpls.dtype, ctx)
    P = ch.nparams
    state = _pack_tree(trafo, optstate, smpls.D) if optstate is not None else np.empty(P, dtype=np.float64)
    if batch_counts is None:
        batch_counts = [e - s for s, e in batch_ranges(smpls.N, nbatches)]
    counts = np.asarray(batch_counts, dtype=np.int64)
    if counts.sum() != smpls.N:
        raise ValueError(f"batch_counts sum to {counts.sum()} but there are {smpls.N} local samples")
    nb = len(counts)
    hist = np.empty(max(nb * nepochs, 1), dtype=np.float64)
    params = np.empty(P, dtype=smpls.dtype)
    n_steps = C.c_int64()
    L.check(ctx._lib.enf_optimize_whitening_batches(
        ch.handle, C.c_void_p(smpls.ptr), nb, counts.ctypes.data_as(C.POINTER(C.c_int64)), int(nepochs),
        float(optimizer.eta), float(optimizer.epsilon),
        L.ENF_NEGLL_ZYGOTE_PRIMAL, 1 if group else 0, 0 if optstate is not None else 1,
        state.ctypes.data_as(C.c_void_p), params.ctypes.data_as(C.c_void_p), hist.ctypes.data_as(C.c_void_p),
        C.byref(n_steps)), ctx.handle)
    ch.last = params.copy()
    result = _rebuild(trafo, _unpack_tree(trafo, params, smpls.D))
    return result, _unpack_tree(trafo, state, smpls.D), [float(v) for v in hist[:n_steps.value]]


def optimize_whitening(smpls, initial_trafo, optimizer: ADAGrad, *, nbatches: int = 100, nepochs: int = 100,
                       optstate=None, negll_history=None, group: bool = False, ctx: Optional[Context] = None,
                       device_loop: bool = False, batch_counts=None):
    """Returns {'result', 'optimizer_state', 'negll_history'} like the
    reference's NamedTuple.  `smpls`: D x N samples (B200Matrix, or a host array
    that is uploaded once).  group=True: `smpls` holds this rank's column block
    of every global batch (see dist.shard_batches); all ranks apply the
    identical update from all-reduced sums.  batch_counts: number of local columns of every batch (default: the
    batches of src/optimize_whitening.jl:31-32 on the LOCAL sample count; with group=True pass
    dist.local_batch_counts(N_global, nbatches, rank, world) whenever the ranks hold different numbers of samples --
    every rank must take the same number of steps).  device_loop=True: the whole loop runs on the device
    (enf_optimize_whitening_batches: two kernel launches per step, no host round trip)."""
    if not isinstance(smpls, B200Matrix):
        X = np.asarray(smpls)
        dt = result_dtype(initial_trafo, X.dtype if X.dtype.kind == "f" else np.float64)
        smpls = B200Matrix.from_host(np.asarray(X, dtype=dt), ctx or default_context())
    if device_loop:
        result, state, hist = _optimize_whitening_device(smpls, copy.deepcopy(initial_trafo), optimizer, nbatches, nepochs,
                                                         optstate, group, batch_counts)
        return {"result": result, "optimizer_state": state, "negll_history": list(negll_history or []) + hist}
    trafo = copy.deepcopy(initial_trafo)
    state = copy.deepcopy(optstate) if optstate is not None else setup(optimizer, trafo)
    hist: List[float] = []
    if batch_counts is None:
        ranges = batch_ranges(smpls.N, nbatches)
    else:
        ends = np.cumsum(np.asarray(batch_counts, dtype=np.int64))
        if len(ends) == 0 or ends[-1] != smpls.N:
            raise ValueError("batch_counts must sum to the number of local samples")
        ranges = [(int(e - c), int(e)) for c, e in zip(batch_counts, ends)]
    for _ in range(nepochs):
        for (s, e) in ranges:
            negll, d_trafo = mvnormal_negll_trafograd(trafo, smpls.cols(s, e), group=group)
            state, trafo = update(optimizer, state, trafo, d_trafo)
            hist.append(float(negll))
    return {"result": trafo, "optimizer_state": state, "negll_history": list(negll_history or []) + hist}
