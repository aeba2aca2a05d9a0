"""Accuracy and speed of the tensor-core second-moment kernel (enf_moments.cu) — scratch tool.
usage: python tools/moments_check.py [D] [N]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import enf_b200 as E
from enf_b200 import _lib as L

D = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
ctx = E.Context(0)
rng = np.random.default_rng(0)
V = rng.normal(size=(D, 16)).astype(np.float32)
f = E.HouseholderTrafo(V)
X = (rng.normal(size=(D, N)) * 1.3 + rng.uniform(-1, 1, (D, 1))).astype(np.float32, order="F")
Xd = E.B200Matrix.from_host(X, ctx)
ch = E.get_chain(f, D, np.float32, ctx)
sums = C.c_void_p()
n = C.c_int64()
L.check(ctx._lib.enf_negll_grad_partial(ch.handle, C.c_void_p(Xd.ptr), N, C.byref(sums), C.byref(n)), ctx.handle)
h = np.empty(n.value, dtype=np.float64)
L.check(ctx._lib.enf_d2h(ctx.handle, h.ctypes.data_as(C.c_void_p), sums, h.nbytes), ctx.handle)
Sh = h.reshape(D + 1, D + 1)
x64 = X.astype(np.float64)
S = x64 @ x64.T
m = x64.sum(1)
scale = np.sqrt((S ** 2).mean())
print(f"D={D} N={N}: max|S-ref|/rms(S) = {np.abs(Sh[:D, :D] - S).max() / scale:.3e}  "
      f"diag rel = {np.abs(np.diag(Sh)[:D] / np.diag(S) - 1).max():.3e}  "
      f"m rel = {np.abs(Sh[:D, D] - m).max() / np.abs(m).max():.3e}  N = {Sh[D, D]}")
# speed (kernel + reduce, resident data)
Nb = int(sys.argv[3]) if len(sys.argv) > 3 else 4_000_000
Xb = E.B200Matrix.randn(D, Nb, np.float32, seed=1, ctx=ctx)
for _ in range(3):
    L.check(ctx._lib.enf_negll_grad_partial(ch.handle, C.c_void_p(Xb.ptr), Nb, None, None), ctx.handle)
ctx.record(0)
for _ in range(5):
    L.check(ctx._lib.enf_negll_grad_partial(ch.handle, C.c_void_p(Xb.ptr), Nb, None, None), ctx.handle)
ctx.record(1)
ms = ctx.elapsed_ms(0, 1) / 5
print(f"moments D={D} N={Nb}: {ms:.3f} ms  {Nb / ms * 1e-6:.3f}e9 samples/s  {Nb * D * 4 / ms * 1e-6:.1f} GB/s  "
      f"{2 * 2.0 * D * D * Nb / ms * 1e-9:.1f} TFLOP/s tf32 issued")
for K in (16, 64):
    fk = E.compose(E.ScaleShiftTrafo(np.full(D, 1.1, np.float32), np.full(D, 0.1, np.float32)),
                   E.HouseholderTrafo(rng.normal(size=(D, K)).astype(np.float32)))
    for nb in (100_000, Nb):
        xb = Xb.cols(0, nb)
        E.mvnormal_negll_trafograd(fk, xb)
        t0 = time.perf_counter()
        for _ in range(5):
            l, g = E.mvnormal_negll_trafograd(fk, xb)
        print(f"enf_negll_grad D={D} K={K} batch={nb}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per call (set_params + moments + chain rule + D2H), negll {l:.6f}")
