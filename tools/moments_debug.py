"""Decode the operand mapping of the moments kernel with a rank-one integer input (scratch tool)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import enf_b200 as E
from enf_b200 import _lib as L

D = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = sys.argv[3] if len(sys.argv) > 3 else "rows"
ctx = E.Context(0)
f = E.HouseholderTrafo(np.random.default_rng(0).normal(size=(D, 16)).astype(np.float32))
if mode == "rows":
    X = np.tile(np.arange(1, D + 1, dtype=np.float32)[:, None], (1, N))     # x[r, j] = r + 1
else:
    X = np.zeros((D, N), np.float32)
    X[int(sys.argv[4]), :] = 1.0
    X[int(sys.argv[5]), :] = 2.0
X = np.asfortranarray(X)
Xd = E.B200Matrix.from_host(X, ctx)
ch = E.get_chain(f, D, np.float32, ctx)
sums = C.c_void_p()
n = C.c_int64()
L.check(ctx._lib.enf_negll_grad_partial(ch.handle, C.c_void_p(Xd.ptr), N, C.byref(sums), C.byref(n)), ctx.handle)
h = np.empty(n.value, dtype=np.float64)
L.check(ctx._lib.enf_d2h(ctx.handle, h.ctypes.data_as(C.c_void_p), sums, h.nbytes), ctx.handle)
Sh = h.reshape(D + 1, D + 1)[:D, :D] / N
ref = (X.astype(np.float64) @ X.astype(np.float64).T) / N
np.set_printoptions(linewidth=250, precision=1, suppress=True)
print("got[:10,:10]\n", Sh[:10, :10])
print("ref[:10,:10]\n", ref[:10, :10])
bad = np.argwhere(np.abs(Sh - ref) > 1e-3 * (np.abs(ref) + 1))
print("mismatches:", len(bad), "of", D * D, "first:", bad[:10].tolist())
if mode == "rows":
    # got[a][b] = (pa+1)(pb+1): recover the row the hardware used for logical a from the first column and the diagonal
    d = np.sqrt(np.maximum(np.diag(Sh), 0))
    print("sqrt(diag) - 1 (row actually used per logical row), first 40:", (d - 1)[:40].round(1).tolist())
    print("rows 120..136:", (d - 1)[120:136].round(1).tolist())
else:
    nz = np.argwhere(np.abs(Sh) > 1e-6)
    print("nonzeros:", [(int(a), int(b), float(Sh[a, b])) for a, b in nz[:20]])
