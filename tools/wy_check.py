"""compact-WY tensor-core kernel vs oracle and vs the dense fold (ENF_NO_WY=1 in a child process is not needed: the SIMT path on an unaligned view is the cross-check)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import enf_b200 as E
from oracle import enf_oracle as O
from chains import both
from conftest import rel_err
ctx = E.default_context()
for spec, D, N in ((["hh64", "ss"], 256, 1000), (["hh64", "ss"], 256, 129), (["ss", "hh32"], 128, 5000), (["hh16", "ss", "hh16"], 128, 2049), (["hh9", "ss"], 256, 100001)):
    fo, fe = both(spec, D, 31, np.float32)
    X = (np.random.default_rng(32).standard_normal((D, N))).astype(np.float32)
    Xd = E.B200Matrix.from_host(X, ctx)
    print(E.get_chain(fe, D, np.float32, ctx).describe()[:60], flush=True)
    t0 = time.time()
    Y, L = E.with_logabsdet_jacobian(fe, Xd)
    y = Y.to_host()
    y_ref, l_ref = O.with_logabsdet_jacobian(fo, X.astype(np.float64))
    print(spec, D, N, "y err", rel_err(y, y_ref), "ladj err", float(np.abs(L.to_host()[0] - l_ref).max()), f"{time.time()-t0:.2f}s", flush=True)
