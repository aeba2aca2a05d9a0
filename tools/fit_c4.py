"""Time the device-side optimize_whitening loop on the C4 chain (D=256, 64 reflections + ScaleShift) — scratch tool."""
import sys
import time

import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import enf_b200 as E
from chains import build

ctx = E.Context(0)
D, nb, bs = 256, 60, 100_000
f = build(E, ["hh64", "ss"], D, np.random.default_rng(44), np.float32)
X = E.B200Matrix.randn(D, nb * bs, np.float32, seed=1, ctx=ctx)
for ne in [int(a) for a in sys.argv[1:]] or (1, 4, 1, 4, 8):
    ctx.sync() if hasattr(ctx, "sync") else None
    t0 = time.perf_counter()
    r = E.optimize_whitening(X, f, E.ADAGrad(), nbatches=nb, nepochs=ne, device_loop=True)
    dt = time.perf_counter() - t0
    print(f"nepochs={ne}: {dt * 1e3:.1f} ms total, {dt / (ne * nb) * 1e6:.1f} us/step, negll {r['negll_history'][0]:.4f} -> {r['negll_history'][-1]:.4f}", flush=True)
