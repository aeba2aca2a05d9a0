"""CPU-only error budget of the gradient path (no GPU needed).

For every gradient test chain: gradients from
  T  = the device algebra (tests/device_model.py) in numpy longdouble (80-bit) -> "truth"
  O  = the oracle (torch float64 autograd of the literal reference formulas)
  M64 = the device algebra in float64, M32 = in float32 (numpy; not MUFU-exact but same conditioning)
and prints, per leaf, err(O vs T), err(M64 vs T), err(M32 vs T) in the parity metric of tests/conftest.py
plus the condition number  kappa = sum_j |integrand_j| / (|sum_j integrand_j| + RMS(leaf))  of the worst row.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import device_model as M
from oracle import enf_oracle as O
from chains import build, flat_grads
from conftest import rel_err


def run_model(leaves, X, dt):
    """forward + backward of the device algebra in dtype dt; returns list of grad dicts (application order)."""
    c = lambda a: np.asarray(a, dtype=dt)
    N = X.shape[1]
    xs = [c(X)]
    for f in leaves:
        n = type(f).__name__
        x = xs[-1]
        if n == "CenterStretch": y, _ = M.cs_fwd(x, c(f.a), c(f.b), c(f.c))
        elif n == "CenterContract": y, _ = M.cc_fwd(x, c(f.a), c(f.b), c(f.c))
        elif n == "JohnsonTrafo": y, _ = M.jo_fwd(x, c(f.gamma), c(f.delta), c(f.xi), c(f.lam))
        elif n == "JohnsonTrafoInv": y, _ = M.ji_fwd(x, c(f.gamma), c(f.delta), c(f.xi), c(f.lam))
        elif n == "ScaleShiftTrafo": y, _ = M.ss_fwd(x, c(f.a), c(f.b))
        else: y, _ = M.hh_fwd(x, c(f.V))
        xs.append(y)
    G = xs[-1].copy()
    out = [None] * len(leaves)
    for i in reversed(range(len(leaves))):
        f = leaves[i]; n = type(f).__name__
        xin, xout = xs[i], xs[i + 1]
        if n == "CenterStretch": G, raw = M.cs_bwd(xin, xout, G, c(f.a), c(f.b), c(f.c)); g = M.cs_finish(raw, N, c(f.a), c(f.b), c(f.c))
        elif n == "CenterContract": G, raw = M.cc_bwd(xin, xout, G, c(f.a), c(f.b), c(f.c)); g = M.cc_finish(raw, N, c(f.a), c(f.b), c(f.c))
        elif n == "JohnsonTrafo": G, raw = M.jo_bwd(xin, xout, G, c(f.gamma), c(f.delta), c(f.xi), c(f.lam)); g = M.jo_finish(raw, N, c(f.gamma), c(f.delta), c(f.xi), c(f.lam))
        elif n == "JohnsonTrafoInv": G, raw = M.ji_bwd(xin, xout, G, c(f.gamma), c(f.delta), c(f.xi), c(f.lam)); g = M.ji_finish(raw, N, c(f.gamma), c(f.delta), c(f.xi), c(f.lam))
        elif n == "ScaleShiftTrafo": G, raw = M.ss_bwd(xin, G, c(f.a), c(f.b)); g = M.ss_finish(raw, N, c(f.a), c(f.b))
        else:
            G, raw, _ = M.hh_bwd(xout, G, c(f.V)); g = M.hh_finish(raw, N, c(f.V))
        out[i] = {k: np.asarray(v, dtype=np.longdouble) / N for k, v in g.items()}
    return out


GRAD_CHAINS = [(["ss", "jo"], 1), (["cc", "jo", "cc", "jo"], 1), (["ss", "hhv", "cc"], 2), (["cc", "jo", "hh4", "ss"], 32),
               (["cs", "ji", "hh3", "ss", "cc", "jo", "hh2", "ss"], 5), (["hh4", "jo", "cs"], 16), (["ji", "hh2", "cs"], 8),
               (["jo", "hh5", "ss"], 100)]

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 3001
    worst = {"O": 0, "M64": 0, "M32": 0}
    for spec, D in GRAD_CHAINS:
        fo = build(O, spec, D, np.random.default_rng(21), np.float64)
        X = (np.random.default_rng(22).standard_normal((D, N)) * 1.2)
        leaves = O.flatten(fo)
        T = run_model(leaves, X, np.longdouble)
        M64 = run_model(leaves, X, np.float64)
        fo32 = build(O, spec, D, np.random.default_rng(21), np.float32)
        X32 = X.astype(np.float32)
        T32 = run_model(O.flatten(fo32), X32.astype(np.float64), np.longdouble)   # truth on the f32-rounded inputs
        M32 = run_model(O.flatten(fo32), X32, np.float32)
        _, g = O.mvnormal_negll_trafograd(fo, X)
        Og = flat_grads(g, fo)
        print(f"--- {spec} D={D} N={N}")
        for i, f in enumerate(leaves):
            for k in T[i]:
                t = np.asarray(T[i][k], dtype=np.float64)
                o = dict(Og)[type(f).__name__ + "." + k] if False else None
            names = [kk for kk in T[i]]
        idx = 0
        for i, f in enumerate(leaves):
            for k in f.fields:
                name, o = Og[idx]; idx += 1
                t = np.asarray(T[i][k]).astype(np.float64).reshape(np.shape(o))
                eo = rel_err(o, t); em = rel_err(np.asarray(M64[i][k]).astype(np.float64).reshape(t.shape), t)
                t32 = np.asarray(T32[i][k]).astype(np.float64).reshape(t.shape)
                e32 = rel_err(np.asarray(M32[i][k]).astype(np.float64).reshape(t.shape), t32)
                worst["O"] = max(worst["O"], eo); worst["M64"] = max(worst["M64"], em); worst["M32"] = max(worst["M32"], e32)
                flag = " <<<" if (em > 1e-12 or e32 > 1e-5) else ""
                print(f"  {name:28s} |g|rms={np.sqrt(np.mean(t*t)):9.3e}  oracle-vs-truth {eo:8.2e}  model64 {em:8.2e}  model32 {e32:8.2e}{flag}")
    print("worst:", worst)
