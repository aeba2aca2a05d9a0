#!/usr/bin/env python
"""C2 (BASELINE configs[1]): 1-D JohnsonTrafo + ScaleShiftTrafo whitening fit via optimize_whitening on 1e7
samples, nbatches=100 (examples/nf_example_1d.jl shape).  Reports time per gradient step."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enf_b200 as E

def main():
    dt = np.float64 if "--f64" in sys.argv else np.float32
    N = 10_000_000
    nep = int(sys.argv[sys.argv.index('--epochs') + 1]) if '--epochs' in sys.argv else 3
    ctx = E.default_context()
    # data: X = (CenterStretch([4],[1],[0]) ∘ JohnsonTrafo([10],[3.5],[10],[1]))(XW), XW ~ N(0,1)   (nf_example_1d.jl:8-15)
    f_true = E.compose(E.CenterStretch(np.array([4.0], dt), np.array([1.0], dt), np.array([0.0], dt)),
                       E.JohnsonTrafo(np.array([10.0], dt), np.array([3.5], dt), np.array([10.0], dt), np.array([1.0], dt)))
    XW = E.B200Matrix.randn(1, N, dt, ctx=ctx)
    X = f_true(XW)
    init = E.compose(E.JohnsonTrafo(np.array([0.0], dt), np.array([5.0], dt), np.array([0.0], dt), np.array([5.0], dt)),
                     E.ScaleShiftTrafo(np.array([1.0], dt), np.array([0.0], dt)))
    E.optimize_whitening(X, init, E.ADAGrad(), nbatches=100, nepochs=1)          # warm
    ctx.sync()
    t = time.perf_counter()
    r = E.optimize_whitening(X, init, E.ADAGrad(), nbatches=100, nepochs=nep)
    ctx.sync()
    dtm = time.perf_counter() - t
    h = r["negll_history"]
    print(f"C2 {np.dtype(dt).name} host loop  : {len(h)} steps in {dtm*1e3:.1f} ms = {dtm/len(h)*1e6:.1f} us/step, "
          f"{N*nep/dtm:.3g} samples/s; negll {h[0]:.4f} -> {h[-1]:.4f}")
    E.optimize_whitening(X, init, E.ADAGrad(), nbatches=100, nepochs=1, device_loop=True)
    t = time.perf_counter()
    r = E.optimize_whitening(X, init, E.ADAGrad(), nbatches=100, nepochs=nep, device_loop=True)
    dtm = time.perf_counter() - t
    h = r["negll_history"]
    print(f"C2 {np.dtype(dt).name} device loop: {len(h)} steps in {dtm*1e3:.1f} ms = {dtm/len(h)*1e6:.1f} us/step, "
          f"{N*nep/dtm:.3g} samples/s; negll {h[0]:.4f} -> {h[-1]:.4f}")

if __name__ == "__main__":
    main()
