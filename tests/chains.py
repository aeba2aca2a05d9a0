"""Chain builders shared by the tests, bench.py and smoke(): the same trafo tree
as an oracle object (oracle/enf_oracle.py) and as a product object (enf_b200)."""
import numpy as np


def rand_params(rng, D, dtype=np.float64):
    """Fixed-seed parameter ranges of SURVEY §8d (C3)."""
    u = lambda lo, hi: rng.uniform(lo, hi, D).astype(dtype)
    return {
        "cs": dict(a=u(0.5, 3), b=u(0.5, 1.5), c=u(-1, 1)),
        "jo": dict(gamma=u(-1, 1), delta=u(1, 3), xi=u(-1, 1), lam=u(0.5, 2)),
        "ss": dict(a=(u(0.5, 2) * rng.choice([-1.0, 1.0], D)).astype(dtype), b=u(-1, 1)),
    }


def build(ns, spec, D, rng, dtype=np.float64):
    """spec: list of op codes in APPLICATION order, e.g. ['hh4','jo','cs'].
    ns: module providing the trafo classes (oracle.enf_oracle or enf_b200).
    Returns the composed trafo (`last ∘ ... ∘ first`)."""
    leaves = []
    for code in spec:
        p = rand_params(rng, D, dtype)
        if code == "cs":
            leaves.append(ns.CenterStretch(**p["cs"]))
        elif code == "cc":
            leaves.append(ns.CenterContract(**p["cs"]))
        elif code == "jo":
            leaves.append(ns.JohnsonTrafo(**p["jo"]))
        elif code == "ji":
            leaves.append(ns.JohnsonTrafoInv(**p["jo"]))
        elif code == "ss":
            leaves.append(ns.ScaleShiftTrafo(**p["ss"]))
        elif code.startswith("hh"):
            K = 1 if code == "hhv" else int(code[2:] or 1)
            V = rng.standard_normal((D, K)).astype(dtype)
            leaves.append(ns.HouseholderTrafo(V[:, 0] if code == "hhv" else V))
        else:
            raise ValueError(code)
    return ns.compose(*reversed(leaves))


def both(spec, D, seed, dtype=np.float64):
    """(oracle_chain, product_chain) with identical parameters."""
    from oracle import enf_oracle as O
    import enf_b200 as E
    return (build(O, spec, D, np.random.default_rng(seed), dtype),
            build(E, spec, D, np.random.default_rng(seed), dtype))


def flat_grads(g, f):
    """Nested gradient dict -> list of (path, array) in application order."""
    if "outer" in g and "inner" in g and hasattr(f, "outer"):
        return flat_grads(g["inner"], f.inner) + flat_grads(g["outer"], f.outer)
    return [(type(f).__name__ + "." + k, np.asarray(v)) for k, v in g.items()]
