"""GPU parity report (run on the B200 box): per-leaf gradient errors at factor 1, in the parity metric of
tests/conftest.py and as plain max-relative error on |ref| > RMS(ref); literal-Float32 comparison of y / ladj.
Prints a table; writes gpurun_out/parity_report.json."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import enf_b200 as E
from oracle import enf_oracle as O
from chains import both, build, flat_grads
from conftest import rel_err


def strict_rel(got, ref):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    rms = np.sqrt(np.mean(ref * ref))
    m = np.abs(ref) > rms
    if not m.any():
        return 0.0
    return float(np.max(np.abs(got - ref)[m] / np.abs(ref)[m]))


GRAD_CHAINS = [(["ss", "jo"], 1), (["cc", "jo", "cc", "jo"], 1), (["ss", "hhv", "cc"], 2), (["cc", "jo", "hh4", "ss"], 32),
               (["cs", "ji", "hh3", "ss", "cc", "jo", "hh2", "ss"], 5), (["hh4", "jo", "cs"], 16), (["ji", "hh2", "cs"], 8),
               (["jo", "hh5", "ss"], 100)]
FWD_CHAINS = [(["cs", "hhv", "ss"], 2), (["jo", "cs"], 1), (["ss", "jo"], 1), (["hh4", "jo", "cs"], 16),
              (["cc", "jo", "hh4", "ss"], 32), (["cs", "jo", "hh4"], 16)]

if __name__ == "__main__":
    ctx = E.default_context()
    out = {"grad": [], "fwd_literal_f32": []}
    for N in (3001, 200_003):
        for dtype in (np.float32, np.float64):
            for spec, D in GRAD_CHAINS:
                if N > 10000 and D > 32:
                    continue
                fo, fe = both(spec, D, 21, dtype)
                X = (np.random.default_rng(22).standard_normal((D, N)) * 1.2).astype(dtype)
                Xd = E.B200Matrix.from_host(X, ctx)
                v_ref, g_ref = O.mvnormal_negll_trafograd(fo, X.astype(np.float64))
                v, g = E.mvnormal_negll_trafograd(fe, Xd)
                worst = (0, "")
                for (k, a), (_, b) in zip(flat_grads(g, fe), flat_grads(g_ref, fo)):
                    b = b.reshape(a.shape)
                    e, s = rel_err(a, b), strict_rel(a, b)
                    out["grad"].append({"N": N, "dtype": np.dtype(dtype).name, "spec": spec, "D": D, "leaf": k, "err": e, "strict": s,
                                        "rms": float(np.sqrt(np.mean(b * b)))})
                    if e > worst[0]:
                        worst = (e, k)
                tol = 1e-5 if dtype == np.float32 else 1e-12
                print(f"N={N:7d} {np.dtype(dtype).name} {str(spec):60s} D={D:3d} negll err {abs(v - v_ref) / (abs(v_ref) + 1):.2e} "
                      f"worst leaf {worst[1]:24s} {worst[0]:.2e} ({worst[0] / tol:.1f} x tol)", flush=True)
    # literal Float32 evaluation of the reference formulas (oracle in float32) vs CUDA f32 vs float64 oracle
    for spec, D in FWD_CHAINS:
        N = 100_000
        fo64 = build(O, spec, D, np.random.default_rng(5), np.float32)   # f32-rounded params
        fe = build(E, spec, D, np.random.default_rng(5), np.float32)
        X = (np.random.default_rng(6).standard_normal((D, N)) * 1.5).astype(np.float32)
        with np.errstate(all="ignore"):
            y32, l32 = O.with_logabsdet_jacobian(fo64, X)                 # all-float32 arithmetic
        assert y32.dtype == np.float32
        y64, l64 = O.with_logabsdet_jacobian(fo64, X.astype(np.float64))
        Y, L = E.with_logabsdet_jacobian(fe, E.B200Matrix.from_host(X, ctx))
        y, l = Y.to_host(), L.to_host()[0]
        fin = np.isfinite(y32).all(0) & np.isfinite(l32)
        r = {"spec": spec, "D": D, "N": N, "literal_f32_nonfinite_cols": int((~fin).sum()),
             "y_cuda_vs_f64": rel_err(y, y64), "y_lit32_vs_f64": rel_err(y32[:, fin], y64[:, fin]), "y_cuda_vs_lit32": rel_err(y[:, fin], y32[:, fin]),
             "l_cuda_vs_f64": rel_err(l, l64), "l_lit32_vs_f64": rel_err(l32[fin], l64[fin]), "l_cuda_vs_lit32": rel_err(l[fin], l32[fin]),
             "y_cuda_vs_f64_strict": strict_rel(y, y64), "l_cuda_vs_f64_strict": strict_rel(l, l64)}
        out["fwd_literal_f32"].append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(out, f, indent=1)
