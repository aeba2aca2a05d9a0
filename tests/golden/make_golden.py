#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/enf_oracle.py).

The reference is pure Julia and Julia is not installed in this image, so these
vectors are NOT outputs of the reference itself; they pin the oracle (which is
pinned in turn to the reference's four known-answer values and its property
tests, tests/test_oracle.py) so that the oracle cannot drift silently, and give
the GPU tests fixed input/output pairs that travel to the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from chains import build, flat_grads  # noqa: E402
from oracle import enf_oracle as O  # noqa: E402

CASES = {
    # name: (spec in application order, D, N, seed)
    "c1_2d_example": (["cs", "hhv", "ss"], 2, 257, 101),          # examples/nf_example_2d.jl:12-15
    "c2_1d_fit": (["ss", "jo"], 1, 513, 102),                      # BASELINE configs[1]
    "c2_1d_example_inv": (["cc", "jo", "cc", "jo"], 1, 300, 103),  # examples/nf_example_1d.jl:19-23
    "c3_d16": (["hh4", "jo", "cs"], 16, 129, 104),                 # BASELINE configs[2]
    "c5_d32": (["cc", "jo", "hh4", "ss"], 32, 130, 105),           # BASELINE configs[4]
    "odd_d5": (["cs", "ji", "hh3", "ss", "cc", "jo", "hh2"], 5, 77, 106),
}


def main():
    for name, (spec, D, N, seed) in CASES.items():
        f = build(O, spec, D, np.random.default_rng(seed))
        X = np.random.default_rng(seed + 1000).standard_normal((D, N)) * 1.3
        Y, ladj = O.with_logabsdet_jacobian(f, X)
        Xi, ladj_i = O.with_logabsdet_jacobian(O.inverse(f), Y)
        negll = float(O.mvnormal_negll_trafo(f, X))
        negll_z, g = O.mvnormal_negll_trafograd(f, X)
        out = {"X": X, "Y": Y, "ladj": ladj, "X_roundtrip": Xi, "ladj_inverse": ladj_i,
               "negll": np.float64(negll), "negll_zygote_primal": np.float64(negll_z)}
        for i, (k, a) in enumerate(flat_grads(g, f)):
            out[f"grad_{i:02d}_{k}"] = a
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "->", {k: getattr(v, "shape", ()) for k, v in list(out.items())[:3]})


if __name__ == "__main__":
    main()
