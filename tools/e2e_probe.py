import os, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import enf_b200 as E
from chains import build
ctx = E.default_context()
D, n = 16, 1 << 25
fe = build(E, ["hh4", "jo", "cs"], D, np.random.default_rng(42), np.float32)
xh = ctx.pinned_empty((D, n), np.float32); yh = ctx.pinned_empty((D, n), np.float32); lh = ctx.pinned_empty((1, n), np.float32)
xh[...] = np.random.default_rng(0).standard_normal((D, 1 << 20)).astype(np.float32).repeat(32, axis=1)
for tag in ("full", "copy_only"):
    if tag == "copy_only": os.environ["ENF_HOST_COPY_ONLY"] = "1"
    E.with_logabsdet_jacobian(fe, xh, out=(yh, lh), ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(3): E.with_logabsdet_jacobian(fe, xh, out=(yh, lh), ctx=ctx)
    dt = (time.perf_counter() - t0) / 3
    print(os.environ.get("ENF_HOST_CHUNK_MB", "32"), tag, f"{n/dt:.4g} samples/s  H2D {n*D*4/dt/1e9:.1f} GB/s  D2H {n*(D+1)*4/dt/1e9:.1f} GB/s")
