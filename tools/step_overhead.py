"""Where does the time of one gradient step go?  kernel (CUDA events) vs C-ABI call (wall) vs Python wrapper (wall)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import enf_b200 as E
from enf_b200 import _lib as L
from chains import build

for spec, D, nb in ((["cc", "jo", "hh4", "ss"], 32, 2_500_000), (["ss", "jo"], 1, 100_000), (["cc", "jo", "hh4", "ss"], 32, 100_000)):
    ctx = E.default_context()
    f = build(E, spec, D, np.random.default_rng(43), np.float32)
    X = E.B200Matrix.randn(D, nb, np.float32, ctx=ctx)
    ch = E.get_chain(f, D, np.float32, ctx)
    lib = ctx._lib
    negll = C.c_double(); g = np.empty(ch.nparams, dtype=np.float32)
    sums, n = C.c_void_p(), C.c_int64()
    for _ in range(5):
        L.check(lib.enf_negll_grad(ch.handle, C.c_void_p(X.ptr), nb, 1, C.byref(negll), g.ctypes.data_as(C.c_void_p)))
    R = 50
    ctx.sync(); ctx.record(0)
    for _ in range(R):
        L.check(lib.enf_negll_grad_partial(ch.handle, C.c_void_p(X.ptr), nb, C.byref(sums), C.byref(n)))
    ctx.record(1); k_ms = ctx.elapsed_ms(0, 1) / R
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(R):
        L.check(lib.enf_negll_grad(ch.handle, C.c_void_p(X.ptr), nb, 1, C.byref(negll), g.ctypes.data_as(C.c_void_p)))
    abi_ms = (time.perf_counter() - t0) / R * 1e3
    t0 = time.perf_counter()
    for _ in range(R):
        E.mvnormal_negll_trafograd(f, X)
    py_ms = (time.perf_counter() - t0) / R * 1e3
    print(f"{spec} D={D} batch={nb}: kernels (grad+reduce, events) {k_ms*1e3:.1f} us | enf_negll_grad wall {abi_ms*1e3:.1f} us | python wrapper wall {py_ms*1e3:.1f} us")
