import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import enf_b200 as E
from chains import both
from oracle import enf_oracle as O
ctx = E.default_context()
for D, spec, N in ((256, ["hh64","ss"], 1000), (256, ["hh16","ss"], 777), (128, ["ss","hh32"], 5000), (64, ["hh8"], 129), (256, ["hh64","ss"], 300001)):
    fo, fe = both(spec, D, 7, np.float32)
    X = np.random.default_rng(8).standard_normal((D, N)).astype(np.float32)
    Xd = E.B200Matrix.from_host(X, ctx)
    Y, L = E.with_logabsdet_jacobian(fe, Xd)
    yr, lr = O.with_logabsdet_jacobian(fo, X.astype(np.float64))
    y = Y.to_host(); l = L.to_host()[0]
    ey = np.max(np.abs(y-yr)/(np.abs(yr)+np.sqrt(np.mean(yr**2))))
    el = np.max(np.abs(l-lr)/(np.abs(lr)+1e-30)) if np.abs(lr).max()>0 else np.abs(l).max()
    print(D, spec, N, "err y %.2e ladj %.2e" % (ey, el), flush=True)
