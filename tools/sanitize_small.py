"""Small invocations of every kernel family for compute-sanitizer (memcheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import enf_b200 as E
from chains import both
ctx = E.default_context()
rng = np.random.default_rng(0)
for spec, D, N, dt in ((["hh64", "ss"], 256, 1300, np.float32), (["ss", "hh32"], 128, 700, np.float32), (["hh4", "jo", "cs"], 16, 5000, np.float32),
                       (["hh4", "jo", "cs"], 16, 3000, np.float64), (["cc", "jo", "hh4", "ss"], 24, 2000, np.float32), (["jo", "ss"], 1, 4001, np.float32),
                       (["cs", "hh2", "jo"], 17, 1001, np.float32)):
    fo, fe = both(spec, D, 3, dt)
    X = E.B200Matrix.from_host(rng.standard_normal((D, N)).astype(dt), ctx)
    Y, L = E.with_logabsdet_jacobian(fe, X)
    X2, L2 = E.with_logabsdet_jacobian(E.inverse(fe), Y)
    v, g = E.mvnormal_negll_trafograd(fe, X)
    ctx.sync()
    print(spec, D, N, np.dtype(dt).name, "ok", float(v), flush=True)
one = np.ones(1, dtype=np.float32)
f2 = E.compose(E.JohnsonTrafo(0 * one, 5 * one, 0 * one, 5 * one), E.ScaleShiftTrafo(one.copy(), 0 * one))
r = E.optimize_whitening(E.B200Matrix.randn(1, 20000, np.float32, ctx=ctx), f2, E.ADAGrad(), nbatches=5, nepochs=4, device_loop=True)
print("fit ok", r["negll_history"][-1])
