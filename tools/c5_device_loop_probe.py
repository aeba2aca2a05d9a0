"""torchrun probe: per-step time of the device-side fit loop on the C5 chain (sharded over the ranks)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
import enf_b200 as E
from chains import build
ctx = E.Context(rank)
if world > 1:
    E.dist.init_group(ctx, rank=rank, world=world)
D, nb = 32, 2_500_000
f = build(E, ["cc", "jo", "hh4", "ss"], D, np.random.default_rng(43), np.float32)
X = E.B200Matrix.randn(D, nb * 8, np.float32, ctx=ctx, col0=rank * nb * 8)
for nep in (1, 5, 20, 20):
    ctx.sync()
    t0 = time.perf_counter()
    r = E.optimize_whitening(X, f, E.ADAGrad(), nbatches=8, nepochs=nep, device_loop=True, group=world > 1)
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"world {world} nepochs {nep:3d}: {dt / (8 * nep) * 1e3:.4f} ms/step  negll {r['negll_history'][0]:.4f} -> {r['negll_history'][-1]:.4f}", flush=True)
