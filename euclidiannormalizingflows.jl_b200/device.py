"""Device context and device-resident sample matrices (host side of the C ABI).

`B200Matrix` is the Python twin of the `B200Matrix{T} <: DenseMatrix{T}` type of
the Julia shim (julia/EuclidianNormalizingFlowsB200.jl): a D x N column-major
matrix that lives in HBM; sample j is the contiguous column j, exactly the
memory of a Julia `Matrix{T}`.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import numpy as np

from . import _lib as L

_DT = {np.dtype(np.float32): L.ENF_F32, np.dtype(np.float64): L.ENF_F64}


def enf_dtype(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _DT:
        raise TypeError(f"trafo chains run in float32 or float64, not {dt}")
    return _DT[dt]


class Context:
    """One CUDA device + stream (enf_ctx).  Not a singleton: one per GPU/process."""

    def __init__(self, device: int = 0):
        self._lib = L.lib()
        h = C.c_void_p()
        L.check(self._lib.enf_init(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)
        self._chains = {}
        self._pinned = []
        self._finalizer = weakref.finalize(self, Context._destroy, self._lib, h, self._chains, self._pinned)

    @staticmethod
    def _destroy(lib, h, chains, pinned):
        for ch in chains.values():
            lib.enf_chain_destroy(ch["handle"])
        chains.clear()
        for p in pinned:
            lib.enf_host_free(h, p)
        pinned.clear()
        lib.enf_destroy(h)

    def close(self):
        self._finalizer()

    def sync(self):
        L.check(self._lib.enf_sync(self.handle), self.handle)

    def record(self, slot: int):
        """Record CUDA event `slot` on the library's stream."""
        L.check(self._lib.enf_event_record(self.handle, int(slot)), self.handle)

    def elapsed_ms(self, a: int, b: int) -> float:
        """Device time between recorded events a and b (waits for b)."""
        ms = C.c_float()
        L.check(self._lib.enf_event_elapsed_ms(self.handle, int(a), int(b), C.byref(ms)), self.handle)
        return float(ms.value)

    @property
    def launches(self) -> int:
        n = C.c_int64()
        L.check(self._lib.enf_launch_count(self.handle, C.byref(n)), self.handle)
        return n.value

    # ---- raw buffers
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        L.check(self._lib.enf_alloc(self.handle, int(nbytes), C.byref(p)), self.handle)
        return p.value

    def free(self, ptr: int):
        L.check(self._lib.enf_free(self.handle, C.c_void_p(ptr)), self.handle)

    def pinned_empty(self, shape, dtype, order="F") -> np.ndarray:
        """numpy array over page-locked host memory (full-rate PCIe copies).
        The memory belongs to the context and is released by Context.close()."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        L.check(self._lib.enf_host_alloc(self.handle, n, C.byref(p)), self.handle)
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        self._pinned.append(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape, order=order)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class B200Matrix:
    """D x N column-major matrix in device memory (or a column-range view of one)."""

    def __init__(self, ctx: Context, D: int, N: int, dtype, _ptr: Optional[int] = None, _owner=None):
        self.ctx = ctx
        self.D, self.N = int(D), int(N)
        self.dtype = np.dtype(dtype)
        enf_dtype(self.dtype)
        self._owner = _owner
        if _ptr is None:
            self.ptr = ctx.alloc(self.nbytes)
            self._fin = weakref.finalize(self, B200Matrix._release, ctx, self.ptr)
        else:
            self.ptr = _ptr
            self._fin = None

    @staticmethod
    def _release(ctx, ptr):
        try:
            ctx.free(ptr)
        except Exception:
            pass

    @property
    def shape(self):
        return (self.D, self.N)

    @property
    def nbytes(self) -> int:
        return self.D * self.N * self.dtype.itemsize

    @classmethod
    def from_host(cls, x: np.ndarray, ctx: Optional[Context] = None) -> "B200Matrix":
        """Upload a (D, N) array (any memory order; it is stored column-major)."""
        ctx = ctx or default_context()
        x = np.asarray(x)
        if x.ndim != 2:
            raise ValueError("expected a D x N matrix")
        xf = np.asfortranarray(x)
        m = cls(ctx, x.shape[0], x.shape[1], xf.dtype)
        if m.nbytes:
            L.check(ctx._lib.enf_h2d(ctx.handle, C.c_void_p(m.ptr), xf.ctypes.data_as(C.c_void_p), m.nbytes), ctx.handle)
            ctx.sync()  # xf may be a temporary
        return m

    @classmethod
    def randn(cls, D: int, N: int, dtype=np.float32, seed: int = 42, col0: int = 0,
              ctx: Optional[Context] = None) -> "B200Matrix":
        """Synthetic N(0,1) samples generated in HBM (enf_fill_normal)."""
        ctx = ctx or default_context()
        m = cls(ctx, D, N, dtype)
        L.check(ctx._lib.enf_fill_normal(ctx.handle, enf_dtype(dtype), C.c_void_p(m.ptr), D, N, col0, seed), ctx.handle)
        return m

    def to_host(self) -> np.ndarray:
        out = np.empty((self.D, self.N), dtype=self.dtype, order="F")
        if self.nbytes:
            L.check(self.ctx._lib.enf_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), self.nbytes),
                    self.ctx.handle)
        return out

    def cols(self, start: int, stop: int) -> "B200Matrix":
        """Contiguous column range [start, stop) as a view: what
        `flatview(batch)` of a partitioned `nestedview(X)` is in the reference
        (src/optimize_whitening.jl:32,38)."""
        start, stop = int(start), int(stop)
        if not (0 <= start <= stop <= self.N):
            raise IndexError((start, stop, self.N))
        return B200Matrix(self.ctx, self.D, stop - start, self.dtype,
                          _ptr=self.ptr + start * self.D * self.dtype.itemsize, _owner=self)

    def astype(self, dtype) -> "B200Matrix":
        """Device-side conversion to the other sample type (enf_convert); self if the type already matches."""
        dtype = np.dtype(dtype)
        if dtype == self.dtype:
            return self
        m = B200Matrix(self.ctx, self.D, self.N, dtype)
        L.check(self.ctx._lib.enf_convert(self.ctx.handle, enf_dtype(dtype), C.c_void_p(m.ptr), enf_dtype(self.dtype),
                                          C.c_void_p(self.ptr), self.D * self.N), self.ctx.handle)
        return m

    def empty_like(self, D: Optional[int] = None) -> "B200Matrix":
        return B200Matrix(self.ctx, self.D if D is None else D, self.N, self.dtype)
